#!/usr/bin/env python3
"""Two more BRIEF pair tables for the table-proof tests (tests/golden/patterns/*.txt).  They are NOT descriptors anyone
should ship: they exist so that a test can rebuild libsvi_gpu.so / the C oracle around a different table and show that
every kernel follows the table (the pair offsets are template immediates in the match kernel).

  alt_random.txt       256 pairs, offsets uniform in [-24, 24] (seed 20261018)
  alt_adversarial.txt  the corner cases of the unrolled code: every offset at +-24 (window corners), pairs of identical
                       points (bit always 0), the same pair repeated, mirrored pairs, all-odd / all-even columns (the
                       odd/even plane split of the match kernel), and a run through (0, 0)."""
import pathlib

import numpy as np

DST = pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "patterns"


def write(name, rows, comment):
    assert len(rows) == 256
    lines = ["# BRIEF-32 test pairs: y1 x1 y2 x2 ; bit = S(y1,x1) < S(y2,x2); test 8j+i -> byte j bit (7-i)", "# " + comment]
    lines += ["%d %d %d %d" % tuple(r) for r in rows]
    (DST / name).write_text("\n".join(lines) + "\n")


def main():
    DST.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(20261018)
    write("alt_random.txt", rng.integers(-24, 25, size=(256, 4)).tolist(), "test table: uniform random offsets (tools/make_alt_patterns.py)")
    rows = []
    c = (-24, 24)
    for y1 in c:                      # 16: all corner-to-corner pairs
        for x1 in c:
            for y2 in c:
                for x2 in c:
                    rows.append((y1, x1, y2, x2))
    rows += [(0, 0, 0, 0), (24, 24, 24, 24), (-24, -24, -24, -24), (5, -7, 5, -7)] * 4      # 16: identical points
    rows += [(3, 4, -5, 6)] * 16                                                              # 16: one pair repeated
    for k in range(16):                                                                        # 32: mirrored pairs
        a = (k - 8, 2 * k - 15, 8 - k, 15 - 2 * k)
        rows += [a, (a[2], a[3], a[0], a[1])]
    odd = np.arange(-23, 24, 2)
    even = np.arange(-24, 25, 2)
    for k in range(48):                                                                        # 48: odd columns only
        rows.append((int(even[k % 25]), int(odd[(5 * k) % 24]), int(odd[(3 * k) % 24]), int(odd[(7 * k + 1) % 24])))
    for k in range(48):                                                                        # 48: even columns only
        rows.append((int(odd[k % 24]), int(even[(5 * k) % 25]), int(even[(3 * k) % 25]), int(even[(7 * k + 1) % 25])))
    for k in range(32):                                                                        # 32: through the centre
        rows.append((0, 0, int(even[k % 25]), int(odd[k % 24])) if k % 2 else (int(odd[k % 24]), int(even[k % 25]), 0, 0))
    rng2 = np.random.default_rng(7)
    while len(rows) < 256:                                                                     # rest: random on the rim
        r = rng2.integers(-24, 25, size=4)
        r[rng2.integers(0, 4)] = 24 if rng2.random() < 0.5 else -24
        rows.append(tuple(int(v) for v in r))
    write("alt_adversarial.txt", rows[:256], "test table: corner cases of the unrolled pair tests (tools/make_alt_patterns.py)")


if __name__ == "__main__":
    main()
