#!/usr/bin/env python3
"""Concurrent host->device bandwidth of one box: N simultaneous pinned copies for several device subsets, plus the PCIe /
NUMA topology (nvidia-smi topo -m).  Explains the e2e scaling of bench.py (every rank streams 0.93 MB per frame from
pinned host memory): prints one JSON document; commit it as profiles/r2_h2d_matrix.json.
   python tools/h2d_matrix.py [--mb 1024] [--reps 4]"""
import argparse
import json
import subprocess

import torch


def measure_duplex(devs, bufs_h, bufs_d, bufs_h2, bufs_d2, reps):
    """H2D and D2H at the same time on every device, 4 : 1 in bytes like the bench's traffic (0.93 MB up, 0.23 MB down per frame)."""
    ev = {}
    for d in devs:
        torch.cuda.synchronize(d)
    for d in devs:
        with torch.cuda.device(d):
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            with torch.cuda.stream(s_up):
                e[0].record()
                for _ in range(reps):
                    bufs_d[d].copy_(bufs_h[d], non_blocking=True)
                e[1].record()
            n4 = bufs_h2[d].numel() // 4
            with torch.cuda.stream(s_dn):
                e[2].record()
                for _ in range(reps):
                    bufs_h2[d][:n4].copy_(bufs_d2[d][:n4], non_blocking=True)
                e[3].record()
            ev[d] = (e, s_up, s_dn, n4)
    up, dn = {}, {}
    for d in devs:
        e, s_up, s_dn, n4 = ev[d]
        s_up.synchronize()
        s_dn.synchronize()
        up[d] = reps * bufs_h[d].numel() / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9
        dn[d] = reps * n4 / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9
    return up, dn


def measure(devs, bufs_h, bufs_d, reps, d2h=False):
    ev = {}
    for d in devs:
        torch.cuda.synchronize(d)
    for d in devs:
        with torch.cuda.device(d):
            s = torch.cuda.Stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                e0.record()
                for _ in range(reps):
                    if d2h:
                        bufs_h[d].copy_(bufs_d[d], non_blocking=True)
                    else:
                        bufs_d[d].copy_(bufs_h[d], non_blocking=True)
                e1.record()
            ev[d] = (e0, e1, s)
    out = {}
    for d in devs:
        ev[d][2].synchronize()
        out[d] = reps * bufs_h[d].numel() / (ev[d][0].elapsed_time(ev[d][1]) * 1e-3) / 1e9
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    n = torch.cuda.device_count()
    bufs_h = {d: torch.empty(a.mb << 20, dtype=torch.uint8, pin_memory=True) for d in range(n)}
    bufs_d = {d: torch.empty(a.mb << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)}
    bufs_h2 = {d: torch.empty(a.mb << 20, dtype=torch.uint8, pin_memory=True) for d in range(n)}
    bufs_d2 = {d: torch.empty(a.mb << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)}
    subsets = [[0]]
    if n >= 2:
        subsets += [[0, 1], [0, n // 2]]
    if n >= 4:
        subsets += [[0, 1, 2, 3], [0, 2, 4, 6][: n // 2] if n >= 8 else [0, 1, 2, 3], [0, 1, n // 2, n // 2 + 1]]
    if n >= 8:
        subsets += [list(range(8))]
    res = []
    for devs in subsets:
        measure(devs, bufs_h, bufs_d, 1)
        r = measure(devs, bufs_h, bufs_d, a.reps)
        r2 = measure(devs, bufs_h, bufs_d, a.reps, d2h=True)
        up, dn = measure_duplex(devs, bufs_h, bufs_d, bufs_h2, bufs_d2, a.reps)
        res.append({"devices": devs, "h2d_gbs_per_device": {str(k): round(v, 2) for k, v in r.items()}, "h2d_gbs_total": round(sum(r.values()), 2),
                    "d2h_gbs_total": round(sum(r2.values()), 2),
                    "duplex_4to1": {"h2d_gbs_total": round(sum(up.values()), 2), "d2h_gbs_total_while_running": round(sum(dn.values()), 2),
                                    "note": "the D2H stream moves a quarter of the bytes and finishes first; h2d_gbs_total is the rate of the H2D stream "
                                            "over its whole duration"}})
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    numa = subprocess.run("lscpu | grep -i -E 'numa|model name|socket|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout
    print(json.dumps({"copy_mb": a.mb, "reps": a.reps, "n_devices": n, "subsets": res, "topology": topo.splitlines(), "cpu": numa.splitlines()}, indent=1))


if __name__ == "__main__":
    main()
