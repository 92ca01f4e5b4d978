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


# ------------------------------------------------------------------ g2o graph files
# What Cg2oOptimizer::optimizeContinuous saves before and after a run (src/optimization/Cg2oOptimizer.cpp:495-514): the
# text graph of g2o (un-vendored third-party library, version "trunk": the tags and field orders below are those of
# g2o's types_slam3d as published -- PARAMS_SE3OFFSET, PARAMS_CAMERAPARAMETERS, VERTEX_SE3:QUAT, VERTEX_TRACKXYZ, FIX,
# EDGE_SE3:QUAT, EDGE_SE3_TRACKXYZ, EDGE_PROJECT_DEPTH, EDGE_PROJECT_DISPARITY -- plus the reference's own
# EDGE_SE3_LINEAR_ACCELERATION, src/optimization/edge_se3_linear_acceleration.cpp:35-112).  The graph is built from the
# front-end's hand-off exactly as the reference builds it:
#   parameters   ids 0..3 = world offset, LEFT camera, RIGHT camera, IMU-to-LEFT offset              (:99-120)
#   landmarks    VERTEX_TRACKXYZ, id = landmark id, optimised position + translation to g2o          (:1143-1152)
#   poses        VERTEX_SE3:QUAT, id = key-frame id + 1000000 (m_uIDShift), the first one fixed       (:41-61, :1229-1240)
#   pose edges   EDGE_SE3:QUAT from the previous key frame, information 100000 * I6 with the translation block scaled by
#                1 / (1 + |t|^2)                                                                         (:1243-1270)
#   gravity      EDGE_SE3_LINEAR_ACCELERATION per key frame, parameter 3, information I3              (:982-997)
#   measurements per key-frame measurement of a landmark in the graph (:1383-1466): skipped unless the squared norm of
#                the landmark estimate seen from the pose is within (0.75, 1.25) of the measured one; factor = 1 / z;
#                |xyz|^2 < 10 -> EDGE_SE3_TRACKXYZ (info factor * 1000 * I3); < 50 -> EDGE_PROJECT_DEPTH (u, v, z; info
#                factor, factor, factor * 100); < 10000 and disparity > 1 -> EDGE_PROJECT_DISPARITY (u, v,
#                disparity / (fx * baseline); info factor, factor, factor * 1000)
# Vertices are written in id order, a FIX line after a fixed vertex, edges in insertion order (g2o's save order);
# numbers with 17 significant digits (lossless).  Robust kernels are not part of the file format.
G2O_ID_SHIFT = 1000000


def _quat_xyzw(R):
    """Eigen::Quaterniond(R) (the branch on the trace / largest diagonal element), normalised as g2o's toVectorQT does."""
    R = np.asarray(R, np.float64)
    t = (R[0, 0] + R[1, 1]) + R[2, 2]
    q = [0.0, 0.0, 0.0, 0.0]   # x y z w
    if t > 0.0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(((R[i, i] - R[j, j]) - R[k, k]) + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    n = np.sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    return [float(v / n) for v in q]


def _num(v) -> str:
    return "%.17g" % float(v)


def _se3(T):
    T = np.asarray(T, np.float64)
    return [T[0, 3], T[1, 3], T[2, 3]] + _quat_xyzw(T[:3, :3])


def _upper(M):
    M = np.asarray(M, np.float64)
    return [M[i, j] for i in range(len(M)) for j in range(i, len(M))]


def _inv_iso(T):
    I = np.eye(4)
    I[:3, :3] = T[:3, :3].T
    for r in range(3):
        I[r, 3] = -((I[r, 0] * T[0, 3] + I[r, 1] * T[1, 3]) + I[r, 2] * T[2, 3])
    return I


def _mul_iso(A, B):
    C = np.eye(4)
    for r in range(3):
        for c in range(4):
            s = 0.0
            for k in range(3):
                s += A[r, k] * B[k, c]
            C[r, c] = s + (A[r, 3] if c == 3 else 0.0)
    return C


def g2o_lines(cam_left, cam_right, baseline_m, keyframes, landmarks, translation_to_g2o=(0.0, 0.0, 0.0)):
    """keyframes: list of dicts {id, T_left_to_world 4x4, acceleration (3,), measurements: [{id, uv_l, uv_r, xyz_left}]};
    landmarks: list of dicts {id, xyz}; cam_*: (fx, fy, cx, cy).  Returns the lines of the graph file."""
    shift = np.asarray(translation_to_g2o, np.float64)
    ident = [0, 0, 0, 0, 0, 0, 1]
    out = ["PARAMS_SE3OFFSET 0 " + " ".join(_num(v) for v in ident),
           "PARAMS_CAMERAPARAMETERS 1 " + " ".join(_num(v) for v in ident + list(cam_left)),
           "PARAMS_CAMERAPARAMETERS 2 " + " ".join(_num(v) for v in ident + list(cam_right)),
           "PARAMS_SE3OFFSET 3 " + " ".join(_num(v) for v in ident)]
    lm = {}
    for l in sorted(landmarks, key=lambda l: l["id"]):
        p = np.asarray(l["xyz"], np.float64) + shift
        lm[int(l["id"])] = p
        out.append("VERTEX_TRACKXYZ %d " % l["id"] + " ".join(_num(v) for v in p))
    poses = []
    for n, kf in enumerate(keyframes):
        T = np.array(kf["T_left_to_world"], np.float64).reshape(4, 4).copy()
        T[:3, 3] += shift
        poses.append(T)
    for n, kf in sorted(enumerate(keyframes), key=lambda e: e[1]["id"]):
        out.append("VERTEX_SE3:QUAT %d " % (kf["id"] + G2O_ID_SHIFT) + " ".join(_num(v) for v in _se3(poses[n])))
        if n == 0:
            out.append("FIX %d" % (kf["id"] + G2O_ID_SHIFT))
    for n, kf in enumerate(keyframes):
        vid = kf["id"] + G2O_ID_SHIFT
        if n > 0:
            M = _mul_iso(_inv_iso(poses[n - 1]), poses[n])
            f = 1.0 / (1.0 + ((M[0, 3] * M[0, 3] + M[1, 3] * M[1, 3]) + M[2, 3] * M[2, 3]))
            info = 100000.0 * np.eye(6)
            info[:3, :3] *= f
            out.append("EDGE_SE3:QUAT %d %d " % (keyframes[n - 1]["id"] + G2O_ID_SHIFT, vid) + " ".join(_num(v) for v in _se3(M) + _upper(info)))
        out.append("EDGE_SE3_LINEAR_ACCELERATION %d 3 " % vid + " ".join(_num(v) for v in list(kf["acceleration"]) + _upper(np.eye(3))))
        Tinv = _inv_iso(poses[n])
        for m in kf["measurements"]:
            if int(m["id"]) not in lm:
                continue
            p, xyz = lm[int(m["id"])], np.asarray(m["xyz_left"], np.float64)
            est = [((Tinv[r, 0] * p[0] + Tinv[r, 1] * p[1]) + Tinv[r, 2] * p[2]) + Tinv[r, 3] for r in range(3)]
            d_abs = (xyz[0] * xyz[0] + xyz[1] * xyz[1]) + xyz[2] * xyz[2]
            d_rel = ((est[0] * est[0] + est[1] * est[1]) + est[2] * est[2]) / d_abs
            if not (0.75 < d_rel < 1.25):
                continue
            f = 1.0 / xyz[2]
            head = "%d %d " % (vid, m["id"])
            if 10.0 > d_abs:
                out.append("EDGE_SE3_TRACKXYZ " + head + "0 " + " ".join(_num(v) for v in list(xyz) + _upper(np.diag([f * 1000, f * 1000, f * 1000]))))
            elif 50.0 > d_abs:
                out.append("EDGE_PROJECT_DEPTH " + head + "1 " + " ".join(_num(v) for v in [np.float32(m["uv_l"][0]), np.float32(m["uv_l"][1]), xyz[2]] + _upper(np.diag([f, f, f * 100]))))
            elif 10000.0 > d_abs:
                # ptUVLEFT.x - ptUVRIGHT.x is a float subtraction (cv::Point2f), widened to double afterwards
                disp = float(np.float32(np.float32(m["uv_l"][0]) - np.float32(m["uv_r"][0])))
                if 1.0 < disp:
                    out.append("EDGE_PROJECT_DISPARITY " + head + "1 " + " ".join(
                        _num(v) for v in [np.float32(m["uv_l"][0]), np.float32(m["uv_l"][1]), disp / (cam_left[0] * baseline_m)] + _upper(np.diag([f, f, f * 1000]))))
    return out


def write_g2o(path, *args, **kw) -> None:
    with open(path, "w") as f:
        f.write("\n".join(g2o_lines(*args, **kw)) + "\n")


def read_g2o(path) -> dict:
    """Tag-wise parse of a g2o text graph: {tag: [list of float fields per line]} with ids kept as ints where they lead."""
    out = {}
    n_ids = {"VERTEX_SE3:QUAT": 1, "VERTEX_TRACKXYZ": 1, "FIX": 1, "PARAMS_SE3OFFSET": 1, "PARAMS_CAMERAPARAMETERS": 1, "EDGE_SE3:QUAT": 2,
             "EDGE_SE3_TRACKXYZ": 3, "EDGE_PROJECT_DEPTH": 3, "EDGE_PROJECT_DISPARITY": 3, "EDGE_SE3_LINEAR_ACCELERATION": 2}
    for line in open(path):
        tok = line.split()
        if not tok:
            continue
        if tok[0] not in n_ids:
            raise ValueError(f"unknown g2o tag {tok[0]}")
        k = n_ids[tok[0]]
        out.setdefault(tok[0], []).append([int(v) for v in tok[1:1 + k]] + [float(v) for v in tok[1 + k:]])
    return out
