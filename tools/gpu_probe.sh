#!/bin/bash
# Probes for the latency paths: parity tests, ncu launch list of one tracking sequence pass (C3), full capture of
# select_corners in the single-pair shape (C1).  usage: bash tools/gpu_probe.sh <tag>
TAG=${1:-probe}; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_tracking.py 2>&1 | tail -1 | cut -c1-600
C3="python bench.py --config c3 --frames 9 --steps 1 --no-cpu-baseline"
if $C3 > $OUT/${TAG}_c3_plain.log 2>&1; then
  tail -1 $OUT/${TAG}_c3_plain.log | cut -c1-300
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'svi' -s 400 -c 120 --csv --log-file $OUT/${TAG}_c3_launches.csv $C3 > $OUT/${TAG}_c3_ncu.log 2>&1
  echo "ncu c3 exit $?"
fi
C1="python bench.py --config c1 --device-only --steps 20"
if $C1 > $OUT/${TAG}_c1_plain.log 2>&1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'select_corners' -s 10 -c 1 -o $OUT/${TAG}_c1_select -f $C1 > $OUT/${TAG}_c1_ncu.log 2>&1
  echo "ncu c1 exit $?"
fi
