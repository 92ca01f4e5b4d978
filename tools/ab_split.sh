#!/bin/bash
# A/B of the matcher knobs on one box (SVI_MATCH_SPLIT = warps per key-point, SVI_MATCH_PRE = LEFT descriptors by the
# pre-pass kernel): parity tests with the defaults, then device-resident values of every combination.
timeout 900 python -m pytest tests -m gpu -x -q -k "stereo_frame_parity or stress_frame or batch or golden or multi_gpu or bounds" 2>&1 | tail -3
for i in 1 2; do
for V in "1 0" "2 0" "1 1" "2 1"; do
  set -- $V
  echo "split $1 pre $2"; SVI_MATCH_SPLIT=$1 SVI_MATCH_PRE=$2 python bench.py --device-only --steps 5 2>&1 | tail -1 | cut -c150-420
done
done
for V in "1 0" "2 1"; do
  set -- $V
  echo "c5 split $1 pre $2"; SVI_MATCH_SPLIT=$1 SVI_MATCH_PRE=$2 python bench.py --config c5 --device-only 2>&1 | tail -1 | cut -c1-420
  echo "c1 split $1 pre $2"; SVI_MATCH_SPLIT=$1 SVI_MATCH_PRE=$2 python bench.py --config c1 --device-only 2>&1 | tail -1 | cut -c1-420
done
