#!/usr/bin/env python3
"""Per-source-line summary of one kernel from an ncu report (needs -lineinfo builds):
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [min_pct]
Prints, per CUDA source line, its share of executed warp instructions and of warp-stall samples."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append((fname, r))
iex, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[iex]) for _, r in lines if r[iex].isdigit())
tots = sum(int(r[ist]) for _, r in lines if r[ist].isdigit())
print(f"total warp instructions {tot}, stall samples {tots}")
for f, r in lines:
    if not r[iex].isdigit():
        continue
    pi, ps = 100 * int(r[iex]) / max(tot, 1), 100 * int(r[ist]) / max(tots, 1) if r[ist].isdigit() else 0
    if pi >= min_pct or ps >= min_pct:
        print(f"{f}:{r[0]:>4} inst {pi:5.1f}%  stall {ps:5.1f}%  {r[1].strip()[:120]}")
