#!/usr/bin/env python3
"""Generate tests/golden/*.npz -- golden vectors for the hot path, made in THIS container where cv2 4.13
is importable (the reference ships no tests or vectors of its own: SURVEY.md section 4).

  stereo_320x240.npz  a small synthetic pair + cv2.cornerHarris response, cv2 GFTT corners (two
                      maxCorners settings, with and without a mask), cv2.BFMatcher results on oracle
                      descriptors, cv2.circle mask, and the oracle's full stereo-frame result.
cv2 runs as the reference runs OpenCV: setUseOptimized(False), setNumThreads(1) (CTrackerGT.cpp:48-49)."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import cv2  # noqa: E402

from oracle import frontend_np as o  # noqa: E402
from svi_mapper_b200 import load_camera  # noqa: E402
from svi_mapper_b200.synth import stereo_pair  # noqa: E402


def main():
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    W, H = 320, 240
    L, R = stereo_pair(W, H, 77, d_max=40)
    resp = cv2.cornerHarris(L, 7, 3, 0.04)
    rng = np.random.default_rng(9)
    centres = np.stack([rng.uniform(0, W, 60), rng.uniform(0, H, 60)], 1).astype(np.float32)
    mask = np.full((H, W), 255, np.uint8)
    for c in centres:
        cv2.circle(mask, (int(np.rint(c[0])), int(np.rint(c[1]))), 7, 0, -1)
    kp300 = o.detect_cv2(L, 300)
    kp300m = o.detect_cv2(L, 300, mask=mask)
    kpall = o.detect_cv2(L, 0)
    # BFMatcher on oracle descriptors
    _, dl = o.brief32(L, kp300.astype(np.float32))
    _, dr = o.brief32(R, kp300.astype(np.float32))
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    m = [bf.match(dl[i:i + 1], dr)[0] for i in range(len(dl))]
    bf_idx = np.array([x.trainIdx for x in m], np.int32)
    bf_dist = np.array([x.distance for x in m], np.float32)
    # full stereo frame by the oracle, cameras = kitti_00 intrinsics on the small frame
    cl = load_camera(str(ROOT / "tests/golden/calib/kitti_00_left.txt"))
    cr = load_camera(str(ROOT / "tests/golden/calib/kitti_00_right.txt"))
    tri = o.Triangulator(o.Camera(W, H, cl.P), o.Camera(W, H, cr.P), o.StereoParams(max_corners=300))
    fr = o.add_new_landmarks(L, R, tri, use_cv2=True)
    out = ROOT / "tests" / "golden" / "stereo_320x240.npz"
    np.savez_compressed(out, left=L, right=R, harris=resp, mask=mask, mask_centres=centres, gftt300=kp300, gftt300_mask=kp300m,
                        gftt_all=kpall, bf_idx=bf_idx, bf_dist=bf_dist, desc_l300=dl, desc_r300=dr,
                        **{"frame_" + k: v for k, v in fr.items()})
    print("wrote", out, out.stat().st_size, "bytes;", len(kp300), len(kp300m), len(kpall), "corners;",
          int((fr["status"] == 0).sum()), "of", len(fr["status"]), "triangulated")


if __name__ == "__main__":
    main()
