#!/usr/bin/env python3
"""C3 measurement: projection-window landmark tracking on a vi_sensor 752x480 pair,
landmarks = frame 0's stereo result with maxCorners raised, next frame = the scene moved by (3, 2) px under a
small claimed camera motion, so the cascade exercises stage 1 (exact projection), stage 2 (window GFTT) and
stage 3 (epipolar line).  Prints one JSON line: GPU ms/frame through the C-ABI (host buffers), stage histogram,
and the CPU restatement (numpy oracle, one thread) timed on a bounded landmark sample with a parity check."""
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from svi_mapper_b200 import StereoFrontend, load_camera  # noqa: E402
from svi_mapper_b200.synth import stereo_pair  # noqa: E402


def main():
    from oracle import frontend_np as o
    calib = ROOT / "tests" / "golden" / "calib"
    cl, cr = load_camera(str(calib / "vi_sensor_left.txt")), load_camera(str(calib / "vi_sensor_right.txt"))
    W, H = cl.width, cl.height
    L, R = stereo_pair(W, H, 4000)
    L1, R1 = np.roll(L, (2, 3), (0, 1)), np.roll(R, (2, 3), (0, 1))
    T = np.eye(4)
    T[0, 3] = 0.01
    with StereoFrontend(cl, cr, max_corners=4000) as fe:
        f0 = fe.add_new_landmarks(L, R)
        ok = np.nonzero(f0["status"] == 0)[0]
        n = len(ok)
        disp = (f0["uv_l"][ok, 0] - f0["uv_r"][ok, 0]).astype(np.float32)
        args = (T, f0["xyz"][ok], f0["desc_l"][ok], f0["desc_r"][ok], disp, 7.0, 1.0)
        kw = dict(uv_reference_left=f0["uv_l"][ok], desc_reference_left=f0["desc_l"][ok], T_left_to_world_at_detection=np.eye(4))
        res = {}
        for name, (a, b) in (("static", (L, R)), ("moved", (L1, R1))):
            for _ in range(3):
                out = fe.track_landmarks(a, b, *args, **kw)
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                out = fe.track_landmarks(a, b, *args, **kw)
            ms = (time.perf_counter() - t0) / reps * 1e3
            res[name] = dict(ms_per_frame=ms, stages=np.bincount(out["stage"], minlength=6).tolist())
        # CPU restatement on a bounded sample of the same landmarks + parity on that sample
        m = min(n, 120)
        tri = o.Triangulator(o.Camera(W, H, cl.P), o.Camera(W, H, cr.P), o.StereoParams(max_corners=4000))
        lms = [dict(xyz_w=f0["xyz"][i], last_desc_l=f0["desc_l"][i], last_desc_r=f0["desc_r"][i], last_disparity=disp[k], size=7.0,
                    uv_ref=f0["uv_l"][i].astype(np.float64), ref_desc_l=f0["desc_l"][i], T_det_l2w=np.eye(4)) for k, i in enumerate(ok[:m])]
        t0 = time.perf_counter()
        ref = o.track_manual_full(L1, R1, tri, T, lms, 1.0)
        cpu_s = time.perf_counter() - t0
        same = all(int(out["stage"][i]) == r["stage"] and int(out["status"][i]) == r["status"] and
                   (not r["stage"] or (np.array_equal(out["xyz"][i], r["xyz"]) and np.array_equal(out["desc_l"][i], r["desc_l"])))
                   for i, r in enumerate(ref))
    line = {"workload": f"vi_sensor {W}x{H}, {n} landmarks, trackManual stages 1-3 (C3)", "gpu": res,
            "gpu_landmarks_per_s_moved": n / (res["moved"]["ms_per_frame"] * 1e-3),
            "cpu_oracle": {"kind": "port (numpy restatement, 1 thread)", "sample_landmarks": m, "seconds": cpu_s,
                           "landmarks_per_s": m / cpu_s, "gpu_matches_cpu_on_sample": bool(same)}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
