#!/usr/bin/env python3
"""Generate svi_mapper_b200/brief_pattern_32.txt -- the 256 BRIEF-32 test pairs.

The genuine table is opencv_contrib modules/xfeatures2d/src/generated_32.i, which is
NOT present in this image (no xfeatures2d, no network) -- SURVEY.md section 8(c) declares
BRIEF "parity unpinned" at pattern level.  This script writes a stand-in with the same
statistics as the BRIEF paper's G-I sampling (isotropic Gaussian, sigma = patch/5 = 9.6,
clipped to the +-24 px patch), with the first byte (8 pairs) set to the values SURVEY.md
recalls for the upstream table.  One line per test: "y1 x1 y2 x2", bit = S(y1,x1) < S(y2,x2).
Dropping the genuine table into the .txt file and re-running tools/gen_pattern_header.py is
the only step needed to switch; no code depends on the values.
"""
import numpy as np, pathlib

FIRST_BYTE = [(-2, -1, 7, -1), (-14, -1, -3, 3), (1, -2, 11, 2), (1, 6, -10, -7),
              (13, 2, -1, 0), (-14, 5, 5, -3), (-2, 8, 2, 4), (-11, 8, -15, 5)]

def main():
    rng = np.random.default_rng(20150531)
    pairs = list(FIRST_BYTE)
    seen = set(pairs)
    while len(pairs) < 256:
        v = np.clip(np.rint(rng.normal(0.0, 48.0 / 5.0, size=4)), -24, 24).astype(int)
        t = tuple(int(a) for a in v)
        if (t[0], t[1]) == (t[2], t[3]) or t in seen:
            continue
        seen.add(t)
        pairs.append(t)
    out = pathlib.Path(__file__).resolve().parents[1] / "svi_mapper_b200" / "brief_pattern_32.txt"
    with open(out, "w") as f:
        f.write("# BRIEF-32 test pairs: y1 x1 y2 x2 ; bit = S(y1,x1) < S(y2,x2); test 8j+i -> byte j bit (7-i)\n")
        f.write("# stand-in table (see tools/make_brief_pattern.py); parity unpinned vs opencv_contrib generated_32.i\n")
        for p in pairs:
            f.write("%d %d %d %d\n" % p)
    print("wrote", out, len(pairs))

if __name__ == "__main__":
    main()
