"""Synthetic stereo SEQUENCES and the per-frame landmark bookkeeping around the GPU tracking call
(BASELINE.json configs[2]: vi_sensor 752x480, frame-to-frame projection-window tracking against ~3000 landmarks).

render_sequence: a static "cardboard cut-out" world -- fronto-parallel textured cards between 2.5 and 25 m in front
of a textured back wall at 40 m (depth 2-40 m) -- rendered through the calibrated P_L / P_R along a
smooth camera trajectory (20 Hz; <= 5 cm and <= 0.5 deg per frame) by exact ray / plane intersection with occlusion,
bilinear texture lookup, +-1 grey-level sensor noise.  Triangulated landmarks therefore reproject consistently from
frame to frame, which is what the projection-window tracker relies on.

SequenceTracker: the image-side loop of CTrackerGT::_trackLandmarks (src/core/CTrackerGT.cpp:137-380) over the
state CFundamentalMatcher keeps per landmark (src/core/CFundamentalMatcher.cpp:1334-2027, 83-193, 2043-2073):
motion scaling, trackManual for every active landmark in ONE call, visibility / failed-tracking bookkeeping, landmark
retirement after 5 failures, re-detection under the mask of the active landmarks.  The tracking / detection calls go
through a backend object (the GPU front-end in the product; tests and bench.py also plug in the CPU restatement to
check the sequence frame by frame)."""
from __future__ import annotations

import numpy as np

from .synth import left_image


# ----------------------------------------------------------------------------- world + rendering
def smooth_trajectory(n_frames: int, seed: int = 0):
    """WORLD->LEFT poses (n, 4, 4): forward motion with lateral sway and a slow yaw / pitch oscillation;
    per frame <= 5 cm translation and <= 0.5 deg rotation, frame 0 = identity."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 4)
    T = np.zeros((n_frames, 4, 4))
    for t in range(n_frames):
        s = t / 20.0                                   # seconds at 20 Hz
        pos = np.array([0.12 * (np.sin(1.7 * s + ph[0]) - np.sin(ph[0])), 0.05 * (np.sin(2.3 * s + ph[1]) - np.sin(ph[1])), 0.5 * s])
        yaw = np.deg2rad(2.0) * (np.sin(1.1 * s + ph[2]) - np.sin(ph[2]))
        pitch = np.deg2rad(1.0) * (np.sin(1.9 * s + ph[3]) - np.sin(ph[3]))
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        R_l2w = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        M = np.eye(4)
        M[:3, :3] = R_l2w.T                            # WORLD->LEFT
        M[:3, 3] = -R_l2w.T @ pos
        T[t] = M
    return T


def make_world(cam_left, seed: int = 0, n_cards: int = 16):
    """Back wall + cards: dicts(z, x0, y0, texel, tex) with world rectangle [x0, x0 + w*texel) x [y0, y0 + h*texel)."""
    f, cx, cy = float(cam_left.P[0, 0]), float(cam_left.P[0, 2]), float(cam_left.P[1, 2])
    W, H = cam_left.width, cam_left.height
    rng = np.random.default_rng(seed)
    planes = []
    z = 40.0
    texel = z / f
    tw, th = int(2.4 * W), int(2.4 * H)
    planes.append(dict(z=z, texel=texel, x0=-(cx / f) * z - 0.7 * W * texel, y0=-(cy / f) * z - 0.7 * H * texel, tex=left_image(tw, th, seed * 100 + 1)))
    depths = np.geomspace(25.0, 2.5, n_cards)
    for i, z in enumerate(depths):
        texel = z / f
        w = int(rng.uniform(0.22, 0.42) * W)
        h = int(rng.uniform(0.25, 0.50) * H)
        u0 = rng.uniform(-0.1 * W, 0.9 * W)            # where the card's corner appears in frame 0
        v0 = rng.uniform(-0.1 * H, 0.8 * H)
        planes.append(dict(z=float(z), texel=texel, x0=(u0 - cx) / f * z, y0=(v0 - cy) / f * z, tex=left_image(w, h, seed * 100 + 2 + i)))
    return planes


def _render(planes, P, baseline_x, T_w2l, W, H, rng):
    f, cx, cy = float(P[0, 0]), float(P[0, 2]), float(P[1, 2])
    R, t = T_w2l[:3, :3], T_w2l[:3, 3]
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    ray = np.stack([(u - cx) / f, (v - cy) / f, np.ones_like(u)], -1)              # camera frame
    ray_w = ray @ R                                                                  # R^T r as row vectors
    centre = np.array([baseline_x, 0.0, 0.0])                                        # camera centre in the LEFT frame
    c_w = (centre - t) @ R                                                           # ... in the world
    img = np.zeros((H, W), np.float64)
    for pl in planes:                                                                # far -> near: nearer cards overwrite
        s = (pl["z"] - c_w[2]) / ray_w[..., 2]
        X = c_w[0] + s * ray_w[..., 0]
        Y = c_w[1] + s * ray_w[..., 1]
        tx = (X - pl["x0"]) / pl["texel"]
        ty = (Y - pl["y0"]) / pl["texel"]
        tex = pl["tex"]
        th, tw = tex.shape
        ok = (s > 0) & (tx >= 0) & (ty >= 0) & (tx < tw - 1) & (ty < th - 1)
        x0 = np.clip(np.floor(tx).astype(np.int64), 0, tw - 2)
        y0 = np.clip(np.floor(ty).astype(np.int64), 0, th - 2)
        ax, ay = tx - x0, ty - y0
        tf = tex.astype(np.float64)
        val = (tf[y0, x0] * (1 - ax) + tf[y0, x0 + 1] * ax) * (1 - ay) + (tf[y0 + 1, x0] * (1 - ax) + tf[y0 + 1, x0 + 1] * ax) * ay
        img = np.where(ok, val, img)
    img = img + rng.integers(-1, 2, size=(H, W))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def render_sequence(cam_left, cam_right, n_frames: int, seed: int = 4000):
    """(L (n, H, W) u8, R (n, H, W) u8, T_world_to_left (n, 4, 4)) for the calibrated rectified pair."""
    W, H = cam_left.width, cam_left.height
    P_l, P_r = np.asarray(cam_left.P, np.float64).reshape(3, 4), np.asarray(cam_right.P, np.float64).reshape(3, 4)
    baseline = -P_r[0, 3] / P_r[0, 0]                  # metres along +x of the LEFT frame
    planes = make_world(cam_left, seed)
    T = smooth_trajectory(n_frames, seed)
    rng = np.random.default_rng(seed + 17)
    L = np.stack([_render(planes, P_l, 0.0, T[t], W, H, rng) for t in range(n_frames)])
    R = np.stack([_render(planes, P_r, baseline, T[t], W, H, rng) for t in range(n_frames)])
    return np.ascontiguousarray(L), np.ascontiguousarray(R), T


# ----------------------------------------------------------------------------- per-frame bookkeeping
def motion_scaling(T_prev_w2l, T_now_w2l) -> float:
    """min(1 + 10*|rotation| + 0.5*|translation|, 5) of the motion since the last frame (CTrackerGT.cpp:157)."""
    M = T_now_w2l @ np.linalg.inv(T_prev_w2l)
    ang = float(np.arccos(np.clip((np.trace(M[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)))
    return float(min(1.0 + 10.0 * ang + 0.5 * float(np.linalg.norm(M[:3, 3])), 5.0))


class SequenceTracker:
    """backend.track(L, R, T_w2l, state: dict of arrays, scaling) -> dict(status, stage, uv_l, uv_r, xyz, desc_l, desc_r)
    backend.add_new(L, R, mask_centres (n, 2) float32) -> dict(uv_l, uv_r, xyz, desc_l, desc_r, status) per key-point."""

    MAX_FAILED = 5            # m_uMaximumFailedSubsequentTrackingsPerLandmark (CFundamentalMatcher.h:83)

    def __init__(self, backend, cam_left, visible_min: int = 100, max_frames_without_detection: int = 2, keypoint_size: float = 7.0):
        self.be, self.cam = backend, cam_left
        self.P = np.asarray(cam_left.P, np.float64).reshape(3, 4)
        self.visible_min, self.max_gap, self.size = visible_min, max_frames_without_detection, np.float32(keypoint_size)
        z = lambda *s, dt=np.float64: np.zeros(s, dt)
        self.s = dict(xyz_w=z(0, 3), last_desc_l=z(0, 32, dt=np.uint8), last_desc_r=z(0, 32, dt=np.uint8), last_disp=z(0, dt=np.float32),
                      uv_ref=z(0, 2), ref_desc_l=z(0, 32, dt=np.uint8), T_det=z(0, 4, 4), failed=z(0, dt=np.int32),
                      visible=z(0, dt=bool), last_uv_l=z(0, 2, dt=np.float32), uid=z(0, dt=np.int64))
        self.next_uid, self.gap, self.visible_last, self.T_prev = 0, 0, 0, None
        self.log = []

    @property
    def n_active(self) -> int:
        return len(self.s["uid"])

    def _mask_centres(self, T_w2l):
        s = self.s
        p = s["xyz_w"] @ T_w2l[:3, :3].T + T_w2l[:3, 3]
        h = p @ self.P[:, :3].T + self.P[:, 3]
        proj = (h[:, :2] / h[:, 2:3]).astype(np.float32)                            # getUV (unrounded) :2062-2065
        return np.where(s["visible"][:, None], s["last_uv_l"], proj).astype(np.float32)

    def process(self, L, R, T_w2l):
        """One frame: trackManual on every active landmark, bookkeeping, re-detection when the trigger fires."""
        s = self.s
        scaling = 1.0 if self.T_prev is None else motion_scaling(self.T_prev, T_w2l)
        self.T_prev = T_w2l.copy()
        stages = np.zeros(6, np.int64)
        n_tracked = self.n_active
        if self.n_active:
            s["visible"][:] = False                                                 # resetVisibilityActiveLandmarks
            r = self.be.track(L, R, T_w2l, s, scaling, self.size)
            hit = r["stage"] > 0
            stages = np.bincount(r["stage"], minlength=6)
            s["last_desc_l"][hit], s["last_desc_r"][hit] = r["desc_l"][hit], r["desc_r"][hit]   # CLandmark::addMeasurement
            s["last_disp"][hit] = r["uv_l"][hit, 0] - r["uv_r"][hit, 0]
            s["last_uv_l"][hit] = r["uv_l"][hit]
            s["failed"][hit] = 0
            s["failed"][~hit] += 1                                                  # :1980-1987, :2000-2001
            s["visible"][:] = hit
            keep = s["failed"] < self.MAX_FAILED                                    # :2005-2009
            for k in s:
                s[k] = s[k][keep]
            self.last_track = r
        n_visible = int(self.s["visible"].sum())
        self.visible_last = n_visible
        n_new = 0
        if self.visible_min > self.visible_last or self.max_gap < self.gap:          # CTrackerGT.cpp:305
            centres = self._mask_centres(T_w2l)
            d = self.be.add_new(L, R, centres)
            ok = d["status"] == 0
            n_new = int(ok.sum())
            T_l2w = np.linalg.inv(T_w2l)
            xyz_w = d["xyz"][ok] @ T_l2w[:3, :3].T + T_l2w[:3, 3]                    # vecPointXYZInitial
            s = self.s
            add = dict(xyz_w=xyz_w, last_desc_l=d["desc_l"][ok], last_desc_r=d["desc_r"][ok],
                       last_disp=(d["uv_l"][ok, 0] - d["uv_r"][ok, 0]).astype(np.float32), uv_ref=d["uv_l"][ok].astype(np.float64),
                       ref_desc_l=d["desc_l"][ok], T_det=np.broadcast_to(T_l2w, (n_new, 4, 4)).copy(), failed=np.zeros(n_new, np.int32),
                       visible=np.ones(n_new, bool), last_uv_l=d["uv_l"][ok].astype(np.float32),
                       uid=np.arange(self.next_uid, self.next_uid + n_new, dtype=np.int64))
            for k in s:
                s[k] = np.concatenate([s[k], add[k]])
            self.next_uid += n_new
            self.visible_last = n_new                                               # :311
            self.gap = 0
            self.last_new = d
        else:
            self.gap += 1
        rec = dict(tracked=int(n_tracked), visible=n_visible, new=n_new, active=self.n_active, scaling=scaling, stages=stages.tolist())
        self.log.append(rec)
        return rec


class GpuBackend:
    """SequenceTracker backend over one StereoFrontend (the product path)."""

    def __init__(self, frontend):
        self.fe = frontend

    def track(self, L, R, T_w2l, s, scaling, size):
        return self.fe.track_landmarks(L, R, T_w2l, s["xyz_w"], s["last_desc_l"], s["last_desc_r"], s["last_disp"], size, scaling,
                                       uv_reference_left=s["uv_ref"], desc_reference_left=s["ref_desc_l"],
                                       T_left_to_world_at_detection=s["T_det"])

    def add_new(self, L, R, centres):
        return self.fe.add_new_landmarks(L, R, mask_centres=centres)
