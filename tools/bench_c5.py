#!/usr/bin/env python3
"""C5 measurement (BASELINE.json configs[4], the stress case): one 3840x1080 pair, maxCorners 10000.
  dense : scan-line search over 1000 px (17 windows of 62 candidates per key-point), through svi_stereo_frames
  sparse: detect both images, describe, match LEFT x RIGHT key-points inside a +-1 row band (svi_match_epipolar)
Prints one JSON line: ms per frame of both modes, counts, and a parity check of the dense mode against the C
restatement of the reference (all host threads; the restatement is the checker and the CPU time beside it)."""
import json
import pathlib
import sys
import time
from types import SimpleNamespace

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from svi_mapper_b200 import StereoFrontend, load_camera  # noqa: E402
from svi_mapper_b200.synth import stereo_pair  # noqa: E402


def main():
    from oracle import c_oracle as co
    calib = ROOT / "tests" / "golden" / "calib"
    k0, k1 = load_camera(str(calib / "kitti_00_left.txt")), load_camera(str(calib / "kitti_00_right.txt"))
    W, H, K, RANGE = 3840, 1080, 10000, 1000.0
    cl, cr = SimpleNamespace(width=W, height=H, P=k0.P), SimpleNamespace(width=W, height=H, P=k1.P)
    L, R = stereo_pair(W, H, 3000, d_max=900)
    out = {"workload": f"{W}x{H}, maxCorners {K}, dense scan-line range {int(RANGE)} px / sparse +-1 row band (C5)"}
    with StereoFrontend(cl, cr, max_corners=K, search_range_px=RANGE, max_candidates=131072, chunk_frames=1) as fe:
        for _ in range(2):
            res = fe.stereo_frames(L[None], R[None])
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            res = fe.stereo_frames(L[None], R[None])
            ts.append(time.perf_counter() - t0)
        got = res.frame(0)
        out["dense"] = dict(ms_per_frame=sorted(ts)[len(ts) // 2] * 1e3, keypoints=int(res.n_keypoints[0]),
                            matched=int((got["status"] == 0).sum()))
        # sparse mode: key-points of both images, descriptors, band matcher
        ts = []
        for it in range(7):
            t0 = time.perf_counter()
            kl, kr = fe.detect(np.stack([L, R]))
            dl, keep_l = fe.describe(L, kl)
            dr, keep_r = fe.describe(R, kr)
            keep_l, keep_r = keep_l.astype(bool), keep_r.astype(bool)
            dl, dr = dl[keep_l], dr[keep_r]
            m = fe.match_epipolar(dl, kl[keep_l], dr, kr[keep_r], 1.0, 1.0, RANGE)
            if it >= 2:
                ts.append(time.perf_counter() - t0)
        out["sparse"] = dict(ms_per_frame=sorted(ts)[len(ts) // 2] * 1e3, queries=int(len(dl)), train=int(len(dr)),
                             matched=int((np.asarray(m[0]) >= 0).sum()))
    cfg = co.make_config(cl, cr, max_corners=K, search_range=RANGE)
    t0 = time.perf_counter()
    ref = co.frame(co.stereo_frames(cfg, L, R, n_threads=1), 0)
    cpu_s = time.perf_counter() - t0
    same = all(np.array_equal(got[k], ref[k]) for k in ("uv_l", "desc_l", "status", "dist", "idx"))
    ok = ref["status"] == 0
    same = same and np.array_equal(got["uv_r"][ok], ref["uv_r"][ok]) and np.array_equal(got["desc_r"][ok], ref["desc_r"][ok])
    out["cpu_port_one_thread"] = dict(seconds_per_frame=cpu_s, gpu_matches_cpu=bool(same))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
