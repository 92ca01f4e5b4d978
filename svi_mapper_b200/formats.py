"""Data formats on either side of the stereo front-end, numpy side.

* key-frame ``.cloud`` files: binary, native endianness, field by field as the reference writes them
  (``src/types/CKeyFrame.cpp:138-186``, read back ``:186-270``): 16 f64 LEFTtoWORLD, u64 point count, per point
  3 f64 world xyz, 3 f64 camera xyz, 4 f64 (uL vL uR vR), u64 descriptor count, that many 32-byte descriptors.
* KITTI odometry poses: 12 numbers per line = rows of the 3x4 LEFTtoWORLD (``src/runnable/tracker_gt.cpp:208-229``).

The C++ twin is ``svi_mapper_b200/host/CKeyFrameCloud.h``; ``tests/test_host.py`` checks that both agree byte for byte.
"""
from __future__ import annotations

import struct

import numpy as np


def write_cloud(path, T_left_to_world, points) -> None:
    """points: iterable of dicts with xyz_world (3,), xyz_camera (3,), uv_l (2,), uv_r (2,), descriptors (n, 32) uint8."""
    points = list(points)
    with open(path, "wb") as f:
        f.write(np.asarray(T_left_to_world, np.float64).reshape(16).tobytes())
        f.write(struct.pack("=Q", len(points)))
        for p in points:
            d = np.ascontiguousarray(np.asarray(p["descriptors"], np.uint8).reshape(-1, 32))
            f.write(np.asarray(p["xyz_world"], np.float64).reshape(3).tobytes())
            f.write(np.asarray(p["xyz_camera"], np.float64).reshape(3).tobytes())
            f.write(np.asarray([p["uv_l"][0], p["uv_l"][1], p["uv_r"][0], p["uv_r"][1]], np.float64).tobytes())
            f.write(struct.pack("=Q", len(d)))
            f.write(d.tobytes())


def read_cloud(path):
    """Returns (T_left_to_world 4x4, list of point dicts); raises ValueError on a truncated or corrupt file."""
    buf = open(path, "rb").read()
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(buf):
            raise ValueError("truncated cloud file")
        out = buf[pos:pos + n]
        pos += n
        return out

    T = np.frombuffer(take(128), np.float64).reshape(4, 4).copy()
    (n,) = struct.unpack("=Q", take(8))
    points = []
    for _ in range(n):
        v = np.frombuffer(take(80), np.float64)
        (nd,) = struct.unpack("=Q", take(8))
        if nd > (1 << 24):
            raise ValueError("corrupt cloud file")
        d = np.frombuffer(take(32 * nd), np.uint8).reshape(nd, 32).copy()
        points.append(dict(xyz_world=v[0:3].copy(), xyz_camera=v[3:6].copy(), uv_l=v[6:8].copy(), uv_r=v[8:10].copy(), descriptors=d))
    if pos != len(buf):
        raise ValueError("trailing bytes in cloud file")
    return T, points


def read_kitti_poses(path) -> np.ndarray:
    """(n, 4, 4) LEFTtoWORLD transforms from a KITTI odometry ground-truth file."""
    rows = []
    for line in open(path):
        if not line.strip():
            continue
        v = np.array(line.split(), np.float64)
        if v.size != 12:
            raise ValueError("malformed pose line")
        T = np.eye(4)
        T[:3, :] = v.reshape(3, 4)
        rows.append(T)
    return np.stack(rows) if rows else np.zeros((0, 4, 4))


def relative_motions(poses: np.ndarray) -> np.ndarray:
    """What tracker_gt.cpp:229 feeds CTrackerGT::process: inverse(T_i) @ T_{i-1}, identity for frame 0."""
    out = np.tile(np.eye(4), (len(poses), 1, 1))
    for i in range(1, len(poses)):
        Ti = np.eye(4)
        Ti[:3, :3] = poses[i][:3, :3].T
        Ti[:3, 3] = -poses[i][:3, :3].T @ poses[i][:3, 3]
        out[i] = Ti @ poses[i - 1]
    return out
