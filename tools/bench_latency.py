#!/usr/bin/env python3
"""Single-pair latency of the drop-in calls a tracker makes once per frame (host buffers in, host buffers out):
svi_stereo_frames with one pair (= addNewLandmarks) at KITTI size and at the VI-sensor size.  Prints one JSON line."""
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
    calib = ROOT / "tests" / "golden" / "calib"
    out = {}
    for name, seed in (("kitti_00", 0), ("vi_sensor", 4000)):
        cl, cr = load_camera(str(calib / f"{name}_left.txt")), load_camera(str(calib / f"{name}_right.txt"))
        L, R = stereo_pair(cl.width, cl.height, seed)
        with StereoFrontend(cl, cr, chunk_frames=1) as fe:
            for _ in range(5):
                res = fe.stereo_frames(L[None], R[None])
            ts = []
            for _ in range(50):
                t0 = time.perf_counter()
                res = fe.stereo_frames(L[None], R[None])
                ts.append(time.perf_counter() - t0)
            ts.sort()
            out[name] = dict(size=[cl.width, cl.height], keypoints=int(res.n_keypoints[0]), ms_median=ts[len(ts) // 2] * 1e3,
                             ms_min=ts[0] * 1e3, ms_p90=ts[int(len(ts) * 0.9)] * 1e3)
    print(json.dumps({"workload": "one stereo pair per call through the host-buffer C-ABI (pageable numpy buffers)", **out}))


if __name__ == "__main__":
    main()
