"""CPU restatement of the per-frame tracker loop around the stereo front-end -- TEST INFRASTRUCTURE ONLY (tests/ compare
the C++ host layer svi_mapper_b200/host/CTrackerGT.h + CFundamentalMatcher.h, which drives the GPU, against this).

What it restates (paths relative to the reference tree):
  CTrackerGT::_trackLandmarks                 src/core/CTrackerGT.cpp:137-380  (motion scaling :157, trackManual :167,
                                              optimizeActiveLandmarks :197, re-detection trigger :305-315)
  CFundamentalMatcher::trackManual            src/core/CFundamentalMatcher.cpp:1334-2027 (candidate rule :1375-1395, image work
                                              by oracle/svi_oracle.c, bookkeeping :1980-2009)
  CFundamentalMatcher::addNewLandmarks        :83-193, getMaskActiveLandmarks :2043-2073
  CLandmark::addMeasurement / optimize        src/types/CLandmark.cpp:80-296, :447-581 (frontend_np.optimize_landmark)
Pose algebra is written out element by element in the order the host layer uses, so that both sides see the same doubles."""
from __future__ import annotations

import numpy as np

from . import c_oracle as co
from . import frontend_np as o

ST_EPI_NO_TRANSLATION = 24
MAX_FAILED = 5


def iso_mul(A, B):
    """Isometry3d * Isometry3d of the host layer: rows 0..2, s = ((a0*b0 + a1*b1) + a2*b2) (+ a3 in the last column)."""
    C = np.eye(4)
    for r in range(3):
        for c in range(4):
            s = 0.0
            for k in range(3):
                s += A[r, k] * B[k, c]
            C[r, c] = s + (A[r, 3] if c == 3 else 0.0)
    return C


def iso_inv(T):
    I = np.eye(4)
    for r in range(3):
        for c in range(3):
            I[r, c] = T[c, r]
    for r in range(3):
        I[r, 3] = -(I[r, 0] * T[0, 3] + I[r, 1] * T[1, 3] + I[r, 2] * T[2, 3])
    return I


def iso_apply(T, p):
    return np.array([T[r, 0] * p[0] + T[r, 1] * p[1] + T[r, 2] * p[2] + T[r, 3] for r in range(3)])


def proj_mul(P, T):
    R = np.zeros((3, 4))
    for r in range(3):
        for c in range(4):
            s = 0.0
            for k in range(4):
                s += P[r, k] * T[k, c]
            R[r, c] = s
    return R


class Landmark:
    def __init__(self, uid, desc_l, desc_r, size, uvl, uvr, xyz_left, T_l2w, P_w2l, P_w2r):
        self.uid, self.ref_desc_l, self.size = uid, desc_l.copy(), float(size)
        self.xyz_opt = iso_apply(T_l2w, xyz_left)
        self.uv_ref = np.array([float(uvl[0]), float(uvl[1])])
        self.failed = self.opt_success = self.opt_failed = 0
        self.optimal = self.visible = False
        self.measurements = []
        self.add_measurement(uvl, uvr, desc_l, desc_r, xyz_left, P_w2l, P_w2r)

    def add_measurement(self, uvl, uvr, desc_l, desc_r, xyz_left, P_w2l, P_w2r):
        self.last_desc_l, self.last_desc_r = desc_l.copy(), desc_r.copy()
        self.last_uv_l = (np.float32(uvl[0]), np.float32(uvl[1]))
        self.last_disp = np.float32(np.float32(uvl[0]) - np.float32(uvr[0]))
        self.last_xyz_left = np.asarray(xyz_left, np.float64).copy()
        self.measurements.append((P_w2l, P_w2r, (np.float32(uvl[0]), np.float32(uvl[1])), (np.float32(uvr[0]), np.float32(uvr[1]))))

    def optimize(self):
        self.optimal = False
        r = o.optimize_landmark(self.xyz_opt, self.measurements)
        self.xyz_opt = np.asarray(r["xyz"], np.float64)
        self.optimal = bool(r["optimal"])
        self.opt_success += r["success"]
        self.opt_failed += r["failed"]


class TrackerGT:
    def __init__(self, cam_l, cam_r, cfg, visible_min=100, max_gap=2, threads=1, native=False):
        self.cam_l, self.cam_r, self.cfg, self.threads, self.native = cam_l, cam_r, cfg, threads, native
        self.P_l, self.P_r = np.asarray(cam_l.P, np.float64).reshape(3, 4), np.asarray(cam_r.P, np.float64).reshape(3, 4)
        self.visible_min, self.max_gap = visible_min, max_gap
        self.points = []            # detection points: dict(T_l2w, landmarks)
        self.visible = []           # m_vecVisibleLandmarks
        self.T_w2l_last = np.eye(4)
        self.frame = self.visible_last = self.detections = self.gap = self.next_uid = 0
        self.tracks = (0, 0, 0)

    def active(self):
        return [lm for d in self.points for lm in d["landmarks"]]

    def process(self, L, R, T_last_to_now, rotation_norm=0.0):
        t = T_last_to_now[:3, 3]
        t_norm = float(np.sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]))
        T_w2l = iso_mul(T_last_to_now, self.T_w2l_last)
        T_l2w = iso_inv(T_w2l)
        scaling = min(1.0 + (10.0 * rotation_norm + 0.5 * t_norm), 5.0)
        for lm in self.visible:
            lm.visible = False
        self.visible = []
        self._track_manual(L, R, T_w2l, T_l2w, scaling)
        self.visible_last = len(self.visible)
        for lm in self.active():
            lm.optimize()
        if self.visible_min > self.visible_last or self.max_gap < self.gap:
            self.visible_last = self._add_new(L, R, T_w2l, T_l2w)
            self.gap = 0
            self.detections += 1
        else:
            self.gap += 1
        self.T_w2l_last = T_w2l
        self.frame += 1

    def _track_manual(self, L, R, T_w2l, T_l2w, scaling):
        P_w2l, P_w2r = proj_mul(self.P_l, T_w2l), proj_mul(self.P_r, T_w2l)
        cand, tdet = [], []
        for d in self.points:
            for lm in d["landmarks"]:
                if 0 < lm.opt_failed:
                    lm.visible = lm.optimal = False
                elif 0 < lm.opt_success and not lm.optimal:
                    lm.visible = False
                else:
                    cand.append(lm)
                    tdet.append(d["T_l2w"])
        s1 = s2 = s3 = 0
        if cand:
            r = co.track_landmarks(self.cfg, L, R, T_w2l, np.array([lm.xyz_opt for lm in cand]), np.array([lm.last_desc_l for lm in cand]),
                                   np.array([lm.last_desc_r for lm in cand]), np.array([lm.last_disp for lm in cand], np.float32),
                                   np.array([lm.size for lm in cand], np.float32), scaling, uv_reference_left=np.array([lm.uv_ref for lm in cand]),
                                   desc_reference_left=np.array([lm.ref_desc_l for lm in cand]), T_left_to_world_at_detection=np.array(tdet),
                                   n_threads=self.threads, native=self.native)
            for i, lm in enumerate(cand):
                if r["stage"][i] > 0:
                    lm.visible, lm.failed = True, 0
                    lm.add_measurement(r["uv_l"][i], r["uv_r"][i], r["desc_l"][i], r["desc_r"][i], r["xyz"][i], P_w2l, P_w2r)
                    self.visible.append(lm)
                    if r["stage"][i] <= 2:
                        s1 += 1
                    elif r["stage"][i] <= 4:
                        s2 += 1
                    else:
                        s3 += 1
                elif r["status"][i] != ST_EPI_NO_TRANSLATION:
                    lm.failed += 1
                    lm.visible = False
        self.tracks = (s1, s2, s3)
        for d in self.points:
            d["landmarks"] = [lm for lm in d["landmarks"] if lm.failed < MAX_FAILED]
        self.points = [d for d in self.points if d["landmarks"]]

    def _mask_centres(self, T_w2l):
        c = []
        for lm in self.active():
            if lm.visible:
                c.append(lm.last_uv_l)
            else:
                p = iso_apply(T_w2l, lm.xyz_opt)
                P = self.P_l
                w = P[2, 0] * p[0] + P[2, 1] * p[1] + P[2, 2] * p[2] + P[2, 3]
                c.append((np.float32((P[0, 0] * p[0] + P[0, 1] * p[1] + P[0, 2] * p[2] + P[0, 3]) / w),
                          np.float32((P[1, 0] * p[0] + P[1, 1] * p[1] + P[1, 2] * p[2] + P[1, 3]) / w)))
        return np.asarray(c, np.float32).reshape(-1, 2)

    def _add_new(self, L, R, T_w2l, T_l2w):
        P_w2l, P_w2r = proj_mul(self.P_l, T_w2l), proj_mul(self.P_r, T_w2l)
        centres = self._mask_centres(T_w2l)
        ok = np.isfinite(centres).all(axis=1) & (np.abs(centres) < 1.0e6).all(axis=1) if len(centres) else np.zeros(0, bool)
        m = o.mask_active_landmarks(L.shape[1], L.shape[0], centres[ok])[None] if len(centres) else None
        d = co.frame(co.stereo_frames(self.cfg, L, R, masks=m, native=self.native), 0)
        new = []
        for u in range(len(d["status"])):
            if d["status"][u] != 0:
                continue
            lm = Landmark(self.next_uid, d["desc_l"][u], d["desc_r"][u], self.cfg.keypoint_size, d["uv_l"][u], d["uv_r"][u], d["xyz"][u], T_l2w, P_w2l, P_w2r)
            lm.optimal = True
            new.append(lm)
            self.next_uid += 1
        if not new:
            return 0
        self.points.append(dict(T_l2w=T_l2w.copy(), landmarks=new))
        return len(new)
