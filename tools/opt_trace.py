#!/usr/bin/env python3
"""Diagnostic for svi_optimize_landmarks: which landmarks of the C3 sequence run into the 1000-iteration cap of
CLandmark::optimize, and what their iteration does (an exact replica of the arithmetic in Python floats).
   python tools/opt_trace.py [n_frames]"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from svi_mapper_b200 import StereoFrontend, load_camera  # noqa: E402
from svi_mapper_b200.sequence import GpuBackend, SequenceTracker, render_sequence  # noqa: E402


def solve(H, b):
    A = [[H[i][j] for j in range(3)] for i in range(4)]
    y = [-b[i] for i in range(4)]
    for c in range(3):
        norm = 0.0
        for r in range(c, 4):
            norm += A[r][c] * A[r][c]
        norm = float(np.sqrt(norm))
        if norm == 0.0:
            continue
        alpha = -norm if A[c][c] > 0.0 else norm
        v = [0.0] * 4
        v[c] = A[c][c] - alpha
        for r in range(c + 1, 4):
            v[r] = A[r][c]
        vv = 0.0
        for r in range(c, 4):
            vv += v[r] * v[r]
        if vv == 0.0:
            continue
        for j in range(c, 3):
            d = 0.0
            for r in range(c, 4):
                d += v[r] * A[r][j]
            for r in range(c, 4):
                A[r][j] -= 2.0 * d / vv * v[r]
        d = 0.0
        for r in range(c, 4):
            d += v[r] * y[r]
        for r in range(c, 4):
            y[r] -= 2.0 * d / vv * v[r]
    x = [0.0] * 3
    for r in (2, 1, 0):
        s2 = y[r]
        for j in range(r + 1, 3):
            s2 -= A[r][j] * x[j]
        x[r] = s2 / A[r][r] if A[r][r] != 0.0 else 0.0
    return x


def trace(x0, ms, cap=1000):
    X = [float(x0[0]), float(x0[1]), float(x0[2]), 1.0]
    prev = 0.0
    hist = []
    for it in range(cap):
        H = [[0.0] * 4 for _ in range(4)]
        b = [0.0] * 4
        tot = 0.0
        for PL, PR, ul, ur in ms:
            J = [[0.0] * 4 for _ in range(4)]
            e = [0.0] * 4
            for s, (P, uv) in enumerate(((PL, ul), (PR, ur))):
                p = [float(v) for v in P]
                a = [p[4 * r] * X[0] + p[4 * r + 1] * X[1] + p[4 * r + 2] * X[2] + p[4 * r + 3] * X[3] for r in range(3)]
                c = a[2]
                e[2 * s] = a[0] / c - float(uv[0])
                e[2 * s + 1] = a[1] / c - float(uv[1])
                for k in range(4):
                    J[2 * s][k] = p[k] / c - a[0] / (c * c) * p[8 + k]
                    J[2 * s + 1][k] = p[4 + k] / c - a[1] / (c * c) * p[8 + k]
            e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]
            w = 10.0 / e2 if 10.0 < e2 else 1.0
            tot += w * e2
            for i in range(4):
                for j in range(4):
                    H[i][j] += w * (J[0][i] * J[0][j] + J[1][i] * J[1][j] + J[2][i] * J[2][j] + J[3][i] * J[3][j])
                b[i] += w * (J[0][i] * e[0] + J[1][i] * e[1] + J[2][i] * e[2] + J[3][i] * e[3])
        dx = solve(H, b)
        for k in range(3):
            X[k] += dx[k]
        hist.append((tot, tuple(X[:3])))
        if 1e-5 > abs(prev - tot):
            return it + 1, hist
        prev = tot
    return cap, hist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    calib = ROOT / "tests" / "golden" / "calib"
    cl, cr = load_camera(str(calib / "vi_sensor_left.txt")), load_camera(str(calib / "vi_sensor_right.txt"))
    P_l, P_r = np.asarray(cl.P, np.float64).reshape(3, 4), np.asarray(cr.P, np.float64).reshape(3, 4)
    L, R, T = render_sequence(cl, cr, n, 4000)
    meas = {}     # uid -> list of (frame, uvl, uvr)
    with StereoFrontend(cl, cr, max_corners=1000) as fe:
        trk = SequenceTracker(GpuBackend(fe), cl)
        for t in range(n):
            uid_before = trk.s["uid"].copy()
            rec = trk.process(L[t], R[t], T[t])
            if len(uid_before):
                r = trk.last_track
                for i in np.nonzero(r["stage"] > 0)[0]:
                    meas.setdefault(int(uid_before[i]), []).append((t, r["uv_l"][i].copy(), r["uv_r"][i].copy()))
            if rec["new"]:
                d = trk.last_new
                ok = np.nonzero(d["status"] == 0)[0]
                for k, i in enumerate(ok):
                    meas.setdefault(int(trk.next_uid - len(ok) + k), []).append((t, d["uv_l"][i].copy(), d["uv_r"][i].copy()))
        PL = np.stack([(P_l @ T[t]).reshape(12) for t in range(n)])
        PR = np.stack([(P_r @ T[t]).reshape(12) for t in range(n)])
        uids = [int(u) for u in trk.s["uid"]]
        guess = trk.s["xyz_w"]
        first, pose, uvl, uvr = [0], [], [], []
        for u in uids:
            for t, a, b in meas.get(u, []):
                pose.append(t); uvl.append(a); uvr.append(b)
            first.append(len(pose))
        import time
        for _ in range(2):
            t0 = time.perf_counter()
            got = fe.optimize_landmarks(guess, first, pose, np.array(uvl), np.array(uvr), PL, PR)
            dt = time.perf_counter() - t0
    it = got["iterations"]
    print(f"{len(uids)} landmarks, {len(pose)} measurements, call {dt * 1e3:.2f} ms; iterations: median {int(np.median(it))}, "
          f"p99 {int(np.percentile(it, 99))}, max {int(it.max())}, capped {(it >= 1000).sum()}, outcomes {np.bincount(got['outcome'], minlength=5).tolist()}")
    for i in np.nonzero(it >= 200)[0][:4]:
        ms = [(PL[t], PR[t], a, b) for t, a, b in meas[uids[i]]]
        k, hist = trace(guess[i], ms)
        print(f"landmark {uids[i]}: {len(ms)} measurements, replica iterations {k} (library {int(it[i])})")
        states = [h[1] for h in hist]
        for p in (1, 2, 3, 4, 6, 8):
            rep = next((j for j in range(p, len(states)) if states[j] == states[j - p]), None)
            print(f"   exact period {p}: first repeat at iteration {rep}")
        for tot, X in hist[-6:]:
            print("   ", repr(tot), [repr(v) for v in X])


if __name__ == "__main__":
    main()
