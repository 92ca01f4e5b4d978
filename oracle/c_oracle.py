"""ctypes wrapper of oracle/svi_oracle.c (TEST INFRASTRUCTURE / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("P_left", C.c_double * 12), ("P_right", C.c_double * 12),
                ("quality_level", C.c_double), ("min_distance", C.c_double), ("harris_k", C.c_double),
                ("min_disparity", C.c_double), ("max_corners", C.c_int32), ("keypoint_size", C.c_float),
                ("search_range", C.c_float), ("match_cutoff", C.c_float)]


class Result(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("uv_left", C.c_void_p), ("uv_right", C.c_void_p), ("xyz_left", C.c_void_p),
                ("desc_left", C.c_void_p), ("desc_right", C.c_void_p), ("distance", C.c_void_p),
                ("match_index", C.c_void_p), ("status", C.c_void_p)]


def build(native: bool = False) -> pathlib.Path:
    target = "libsvi_oracle_native.so" if native else "libsvi_oracle.so"
    r = subprocess.run(["make", "-C", str(HERE), "native" if native else "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return HERE / target


_libs: dict = {}


def load(native: bool = False):
    """Load (building if needed and possible) the oracle library."""
    key = bool(native)
    if key in _libs:
        return _libs[key]
    path = HERE / ("libsvi_oracle_native.so" if native else "libsvi_oracle.so")
    src = HERE / "svi_oracle.c"
    if not path.exists() or path.stat().st_mtime < src.stat().st_mtime:
        path = build(native)
    lib = C.CDLL(str(path))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    lib.svo_harris_response.argtypes = [vp, ci, ci, ci, C.c_double, vp]
    lib.svo_gftt.argtypes = [vp, ci, ci, ci, vp, ci, ci, C.c_double, C.c_double, C.c_double, vp, ci]
    lib.svo_brief32.argtypes = [vp, ci, ci, ci, vp, ci, vp, vp]
    lib.svo_match.argtypes = [vp, vp, ci, vp]
    lib.svo_point_in_left.argtypes = [C.POINTER(Config), vp, vp, vp]
    lib.svo_triangulate_right.argtypes = [C.POINTER(Config), vp, ci, cf, cf, cf, vp, vp, vp, vp, vp, vp, vp]
    lib.svo_triangulate_left.argtypes = [C.POINTER(Config), vp, ci, cf, cf, cf, cf, vp, vp, vp, vp, vp, vp, vp]
    lib.svo_stereo_frames_mt.argtypes = [C.POINTER(Config), vp, vp, ci, C.c_size_t, ci, vp, C.POINTER(Result), vp, vp, ci]
    _libs[key] = lib
    return lib


def make_config(cam_left, cam_right, max_corners=1000, quality_level=0.01, min_distance=7.0, harris_k=0.04,
                min_disparity=0.01, keypoint_size=7.0, search_range=60.0, match_cutoff=100.0) -> Config:
    c = Config()
    c.width, c.height = int(cam_left.width), int(cam_left.height)
    pl, pr = np.asarray(cam_left.P, np.float64).reshape(12), np.asarray(cam_right.P, np.float64).reshape(12)
    for i in range(12):
        c.P_left[i], c.P_right[i] = float(pl[i]), float(pr[i])
    c.quality_level, c.min_distance, c.harris_k, c.min_disparity = quality_level, min_distance, harris_k, min_disparity
    c.max_corners, c.keypoint_size, c.search_range, c.match_cutoff = max_corners, keypoint_size, search_range, match_cutoff
    return c


def harris_response(img: np.ndarray, k=0.04, native=False) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    rc = load(native).svo_harris_response(img.ctypes.data, w, h, w, k, out.ctypes.data)
    assert rc == 0
    return out


def gftt(img: np.ndarray, max_corners=1000, quality=0.01, min_distance=7.0, mask=None, k=0.04, native=False) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = max_corners if max_corners > 0 else w * h // 4
    xy = np.zeros((cap, 2), np.int32)
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    n = load(native).svo_gftt(img.ctypes.data, w, h, w, m.ctypes.data if m is not None else None, w, max_corners, quality,
                              min_distance, k, xy.ctypes.data, cap)
    assert n >= 0
    return xy[:n].copy()


def brief32(img: np.ndarray, pts, native=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 2))
    n = len(p)
    desc = np.zeros((max(n, 1), 32), np.uint8)
    kept = np.zeros(max(n, 1), np.int32)
    nk = load(native).svo_brief32(img.ctypes.data, w, h, w, p.ctypes.data, n, desc.ctypes.data, kept.ctypes.data)
    assert nk >= 0
    return kept[:nk].copy(), desc[:nk].copy()


def stereo_frames(cfg: Config, left: np.ndarray, right: np.ndarray, masks=None, n_threads: int = 1, native=False) -> dict:
    """svo_stereo_frames_mt on (n, H, W) uint8 batches -> dict of (n, cap, ...) arrays + counts."""
    L = np.ascontiguousarray(left, np.uint8)
    R = np.ascontiguousarray(right, np.uint8)
    if L.ndim == 2:
        L, R = L[None], R[None]
    n, h, w = L.shape
    cap = int(cfg.max_corners)
    M = np.ascontiguousarray(masks, np.uint8).reshape(n, h, w) if masks is not None else None
    out = dict(uv_left=np.zeros((n, cap, 2), np.float32), uv_right=np.zeros((n, cap, 2), np.float32),
               xyz_left=np.zeros((n, cap, 3), np.float64), desc_left=np.zeros((n, cap, 32), np.uint8),
               desc_right=np.zeros((n, cap, 32), np.uint8), distance=np.full((n, cap), -1, np.int32),
               match_index=np.full((n, cap), -1, np.int32), status=np.zeros((n, cap), np.uint8),
               n_keypoints=np.zeros(n, np.int32), n_detected=np.zeros(n, np.int32))
    r = Result(cap, out["uv_left"].ctypes.data, out["uv_right"].ctypes.data, out["xyz_left"].ctypes.data,
               out["desc_left"].ctypes.data, out["desc_right"].ctypes.data, out["distance"].ctypes.data,
               out["match_index"].ctypes.data, out["status"].ctypes.data)
    rc = load(native).svo_stereo_frames_mt(C.byref(cfg), L.ctypes.data, R.ctypes.data, w, w * h, n,
                                           M.ctypes.data if M is not None else None, C.byref(r),
                                           out["n_keypoints"].ctypes.data, out["n_detected"].ctypes.data, int(n_threads))
    assert rc == 0
    return out


def frame(out: dict, f: int) -> dict:
    n = int(out["n_keypoints"][f])
    return dict(uv_l=out["uv_left"][f, :n], uv_r=out["uv_right"][f, :n], xyz=out["xyz_left"][f, :n],
                desc_l=out["desc_left"][f, :n], desc_r=out["desc_right"][f, :n], dist=out["distance"][f, :n],
                idx=out["match_index"][f, :n], status=out["status"][f, :n])


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
