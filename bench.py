#!/usr/bin/env python3
"""bench.py -- stereo frames/sec of the new-landmark hot path (detect + describe + epipolar match +
triangulate, = CFundamentalMatcher::addNewLandmarks per pair) on synthetic KITTI-00-shaped data.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one batch of --frames (default 4096) stereo pairs of
1241x376 with maxCorners 2000 (BASELINE.json configs[1]); every rank owns its own batch (frames
are independent: weak scaling, no collective on the data path).  Prints ONE JSON line on rank 0.

  value      whole-job frames/s with inputs and outputs resident in HBM (svi_stereo_frames_device),
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks
  e2e        the same metric through the host-buffer C-ABI call (svi_stereo_frames): pinned host
             images -> device -> results back in pinned host memory, every step
  roofline   dominant kernel's algorithmic bytes per launch / its mean launch duration (CUDA events on
             the library's own streams inside the timed region) against the measured HBM peak
  cpu_baseline  the C restatement of the reference CPU path (oracle/svi_oracle.c) timed on this box's
             host cores on a bounded sample of the same frames, and used as the parity checker
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 1241, 376
MAX_CORNERS = 2000
BYTES_PER_KEYPOINT = 113  # uvL 8 + uvR 8 + xyz 24 + descL 32 + descR 32 + dist 4 + idx 4 + status 1
CALIB = ROOT / "tests" / "golden" / "calib"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="stereo pairs per rank per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="tuning aid: only the device-resident measurement (no e2e, no CPU baseline); not a bench line")
    return ap.parse_args()


def cameras():
    from svi_mapper_b200 import load_camera
    return load_camera(str(CALIB / "kitti_00_left.txt")), load_camera(str(CALIB / "kitti_00_right.txt"))


def algorithmic_bytes_per_frame() -> int:
    """every input byte read once + 113 B per key-point written once (K = maxCorners)."""
    return 2 * W * H + BYTES_PER_KEYPOINT * MAX_CORNERS


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_fps(frames_l, frames_r, cams, target_seconds: float, threads: int):
    """Time the C restatement of the reference CPU path on a bounded sample; returns
    (fps_all_threads, fps_one_thread, n_sample, result_of_sample)."""
    from oracle import c_oracle as co
    native = True
    try:
        co.load(native=True)
    except Exception:
        native = False
    cfg = co.make_config(cams[0], cams[1], max_corners=MAX_CORNERS)
    n_avail = frames_l.shape[0]
    t0 = time.perf_counter()
    co.stereo_frames(cfg, frames_l[:1], frames_r[:1], n_threads=1, native=native)
    t_one = time.perf_counter() - t0
    n = int(max(threads, min(n_avail, target_seconds * threads / max(t_one, 1e-3))))
    n = max(1, min(n, n_avail))
    t0 = time.perf_counter()
    out = co.stereo_frames(cfg, frames_l[:n], frames_r[:n], n_threads=threads, native=native)
    t_all = time.perf_counter() - t0
    return n / t_all, 1.0 / t_one, n, out, native


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU algorithm (C restatement; the reference binary cannot be
    built here) on the host cores, every step a bounded sample of the same workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import c_oracle as co
    from svi_mapper_b200.synth import stereo_pair
    cams = cameras()
    threads = co.host_threads()
    native = True
    try:
        co.load(native=True)
    except Exception:
        native = False
    cfg = co.make_config(cams[0], cams[1], max_corners=MAX_CORNERS)
    n_unique = min(8, max(2, threads))
    pairs = [stereo_pair(W, H, 1000 + i) for i in range(n_unique)]
    sample = max(threads, 8)
    L = np.stack([pairs[i % n_unique][0] for i in range(sample)])
    R = np.stack([pairs[i % n_unique][1] for i in range(sample)])
    for _ in range(min(args.warmup, 1)):
        co.stereo_frames(cfg, L, R, n_threads=threads, native=native)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.stereo_frames(cfg, L, R, n_threads=threads, native=native)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "stereo frames/sec (detect+describe+match+triangulate) at 1241x376",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64 (Harris/triangulation), u8/u16/u32 (BRIEF, Hamming)", "data": "synthetic",
        "config": {"workload": f"KITTI-00-shaped synthetic 1241x376 pairs, maxCorners {MAX_CORNERS}, 60 px scan-line range; "
                               f"bounded sample of {sample} frames per step on the host CPU"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frames/step x {args.steps} steps, {threads} threads, "
                                   f"{'-march=native' if native else 'x86-64-v3'} build of oracle/svi_oracle.c"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    import numpy as np
    import torch
    import torch.distributed as dist

    from svi_mapper_b200 import StereoFrontend, _lib
    from svi_mapper_b200.synth import stereo_batch_torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cams = cameras()
    F, cap = args.frames, MAX_CORNERS
    fe = StereoFrontend(cams[0], cams[1], device=local_rank, max_corners=MAX_CORNERS)
    cfg = fe.config()

    # ---- synthetic batch, generated on the device; every frame distinct
    dL, dR = stereo_batch_torch(F, W, H, seed=1000 + rank, device=dev)
    torch.cuda.synchronize()

    def dev_outputs():
        t = dict(n_kp=torch.zeros(F, dtype=torch.int32, device=dev), n_det=torch.zeros(F, dtype=torch.int32, device=dev),
                 uv_l=torch.zeros(F, cap, 2, device=dev), uv_r=torch.zeros(F, cap, 2, device=dev),
                 xyz=torch.zeros(F, cap, 3, dtype=torch.float64, device=dev),
                 dl=torch.zeros(F, cap, 32, dtype=torch.uint8, device=dev), dr=torch.zeros(F, cap, 32, dtype=torch.uint8, device=dev),
                 dist=torch.zeros(F, cap, dtype=torch.int32, device=dev), idx=torch.zeros(F, cap, dtype=torch.int32, device=dev),
                 st=torch.zeros(F, cap, dtype=torch.uint8, device=dev))
        r = _lib.StereoResult(cap, t["n_kp"].data_ptr(), t["n_det"].data_ptr(), t["uv_l"].data_ptr(), t["uv_r"].data_ptr(),
                              t["xyz"].data_ptr(), t["dl"].data_ptr(), t["dr"].data_ptr(), t["dist"].data_ptr(),
                              t["idx"].data_ptr(), t["st"].data_ptr())
        return t, r

    dout, dres = dev_outputs()
    stream = torch.cuda.current_stream().cuda_stream

    def device_step():
        fe.stereo_frames_device(dL.data_ptr(), dR.data_ptr(), W, W * H, F, dres, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        device_step()
    barrier()
    fe.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    fe.check_overflow()   # the device-resident entry point reports candidate-list overflow here (never a silent cut)
    ms = e0.elapsed_time(e1)
    stages = fe.stage_timings()
    fe.set_profiling(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * F * args.steps / (ms_max * 1e-3)

    if args.device_only:
        if rank == 0:
            print(json.dumps({"device_only": True, "value": value, "ms_per_step": ms_max / args.steps, "clocks": clocks,
                              "stage_ms": {k: v["total_ms"] for k, v in stages.items()}}), flush=True)
        fe.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the host-buffer C-ABI call: pinned host in, pinned host out, every step
    hL = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True).copy_(dL)
    hR = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True).copy_(dR)
    ho = dict(n_kp=torch.zeros(F, dtype=torch.int32, pin_memory=True), n_det=torch.zeros(F, dtype=torch.int32, pin_memory=True),
              uv_l=torch.zeros(F, cap, 2, pin_memory=True), uv_r=torch.zeros(F, cap, 2, pin_memory=True),
              xyz=torch.zeros(F, cap, 3, dtype=torch.float64, pin_memory=True),
              dl=torch.zeros(F, cap, 32, dtype=torch.uint8, pin_memory=True), dr=torch.zeros(F, cap, 32, dtype=torch.uint8, pin_memory=True),
              dist=torch.zeros(F, cap, dtype=torch.int32, pin_memory=True), idx=torch.zeros(F, cap, dtype=torch.int32, pin_memory=True),
              st=torch.zeros(F, cap, dtype=torch.uint8, pin_memory=True))
    hres = _lib.StereoResult(cap, ho["n_kp"].data_ptr(), ho["n_det"].data_ptr(), ho["uv_l"].data_ptr(), ho["uv_r"].data_ptr(),
                             ho["xyz"].data_ptr(), ho["dl"].data_ptr(), ho["dr"].data_ptr(), ho["dist"].data_ptr(),
                             ho["idx"].data_ptr(), ho["st"].data_ptr())
    torch.cuda.synchronize()

    def host_step():
        fe.stereo_frames_raw(hL.data_ptr(), hR.data_ptr(), W, W * H, F, hres)

    for _ in range(max(1, args.warmup // 2)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * F * args.steps / float(t_e.item())
    h2d = 2 * F * W * H
    d2h = F * cap * BYTES_PER_KEYPOINT + F * 8
    # what the host link can do on this box: one plain pinned -> device copy of the LEFT batch, best of 3
    link_gbs = 0.0
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        barrier()
        c0.record()
        dL.copy_(hL, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link_gbs = max(link_gbs, hL.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9)

    # the two entry points must agree with each other on the whole batch
    torch.cuda.synchronize()
    same = bool(torch.equal(dout["n_kp"].cpu(), ho["n_kp"]))
    nk = ho["n_kp"].numpy()
    total_kp = int(nk.sum())
    total_ok = int((ho["st"].numpy() == 0)[np.arange(cap)[None, :] < nk[:, None]].sum())

    if rank == 0:
        n_chunks = (F + cfg["chunk_frames"] - 1) // cfg["chunk_frames"]
        launches = args.steps * n_chunks * 4   # harris_box, boxsum9, select_corners, stereo_match per chunk
        dom = max(stages, key=lambda k: stages[k]["total_ms"]) if stages else None
        peak, peak_src = measured_hbm_peak()
        roofline = None
        if dom and stages[dom]["launches"]:
            avg_ms = stages[dom]["total_ms"] / stages[dom]["launches"]
            frames_per_launch = min(cfg["chunk_frames"], F)
            achieved = frames_per_launch * algorithmic_bytes_per_frame() / (avg_ms * 1e-3) / 1e9
            traffic = None
            tf = ROOT / "profiles" / "roofline_traffic.json"
            if tf.exists():
                try:
                    traffic = json.loads(tf.read_text()).get(dom, {}).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            # Six lanes run concurrently, so the CUDA-event bracket of one launch also contains the time the GPU spent on
            # other lanes' kernels (the per-stage event sums add up to several times the step).  The launch duration
            # the roofline uses is therefore the step time attributed to the stage by its share of those event sums:
            # step_ms * share / launches_per_step -- it agrees with the isolated ncu duration of the kernel
            # (profiles/roofline_traffic.json), the raw bracket is kept as avg_launch_ms_concurrent.
            tot_ms = sum(v["total_ms"] for v in stages.values())
            share = stages[dom]["total_ms"] / tot_ms
            excl_ms = (ms_max / args.steps) * share / (stages[dom]["launches"] / args.steps)
            achieved = frames_per_launch * algorithmic_bytes_per_frame() / (excl_ms * 1e-3) / 1e9
            path_gbs = algorithmic_bytes_per_frame() * (value / world) / 1e9
            ncu_us = None
            try:
                ncu_us = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text()).get(dom, {}).get("ncu_us_per_launch")
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "avg_launch_ms": excl_ms, "avg_launch_ms_concurrent": avg_ms, "step_share": share,
                        "ncu_isolated_launch_us": ncu_us, "frames_per_launch": frames_per_launch,
                        "algorithmic_bytes_per_frame": algorithmic_bytes_per_frame(),
                        "whole_path": {"achieved": path_gbs, "frac": path_gbs / peak,
                                       "note": "SURVEY.md 8d: B_frame x frames/s per GPU over the measured HBM peak"},
                        "limiter": "L1/shared-memory data pipe and instruction issue (ncu: l1tex data pipe 63-88 % of peak in "
                                   "harris_box / boxsum9 / stereo_match, DRAM < 3 % of peak in every kernel), not HBM",
                        "stage_share": {k: v["total_ms"] for k, v in stages.items()}}
        line = {
            "metric": "stereo frames/sec (detect+describe+match+triangulate) at 1241x376",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64 (Harris/triangulation), u8/u16/u32 (BRIEF, Hamming)", "data": "synthetic",
            "config": {"workload": f"KITTI-00-shaped synthetic batch of {F} stereo pairs per GPU at {W}x{H}, maxCorners "
                                   f"{MAX_CORNERS}, 60 px scan-line range (BASELINE.json configs[1])",
                       "frames_per_gpu": F, "chunk_frames": cfg["chunk_frames"], "lanes": cfg["n_lanes"],
                       "l2_policy": f"inputs {2 * F * W * H / 1e9:.2f} GB per step >> 126 MB L2 (no flush needed)",
                       "keypoints_per_frame": total_kp / F, "matched_per_frame": total_ok / F,
                       "device_vs_host_entry_equal": same},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "h2d_gbs_achieved": h2d * (e2e_value / world / F) / 1e9, "h2d_gbs_plain_copy": link_gbs,
                    "note": "host link bound when h2d_gbs_achieved is close to h2d_gbs_plain_copy (rank 0's link)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle as co
            threads = co.host_threads()
            n_cpu = min(F, 4 * threads + 8)
            sl = hL[:n_cpu].numpy()
            sr = hR[:n_cpu].numpy()
            fps_all, fps_one, n_s, ref, native = cpu_reference_fps(sl, sr, cams, args.cpu_seconds, threads)
            # parity gate on the sampled frames: GPU (host entry) vs the CPU restatement
            ok = True
            for f in range(n_s):
                k = int(ref["n_keypoints"][f])
                ok &= k == int(nk[f])
                if not ok:
                    break
                for a, b in (("uv_left", "uv_l"), ("desc_left", "dl"), ("status", "st"), ("distance", "dist"), ("match_index", "idx")):
                    ok &= bool(np.array_equal(ref[a][f, :k], ho[b][f, :k].numpy()))
                good = ref["status"][f, :k] == 0
                ok &= bool(np.array_equal(ref["uv_right"][f, :k][good], ho["uv_r"][f, :k].numpy()[good]))
                ok &= bool(np.array_equal(ref["desc_right"][f, :k][good], ho["dr"][f, :k].numpy()[good]))
                ok &= bool(np.allclose(ref["xyz_left"][f, :k][good], ho["xyz"][f, :k].numpy()[good], rtol=1e-5, atol=0))
            line["cpu_baseline"] = {"value": fps_all, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "single_thread_value": fps_one,
                                    "sample": f"first {n_s} frames of the batch, {threads} threads (one frame per thread), "
                                              f"{'-march=native' if native else 'x86-64-v3'} build of oracle/svi_oracle.c",
                                    "gpu_matches_cpu_on_sample": bool(ok)}
        print(json.dumps(line), flush=True)
    fe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
