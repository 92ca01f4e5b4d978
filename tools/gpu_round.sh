#!/bin/bash
# One GPU-box session: parity tests, the default bench line, every BASELINE configuration, the ncu launch list and one
# --set full capture of the four kernels of the new-landmark path.  Usage (repo root, under gpurun): bash tools/gpu_round.sh <tag> [quick]
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench_c2.err
echo "bench c2 exit $?"; cut -c1-400 $OUT/${TAG}_bench_c2.json
if [ "${2:-}" != "quick" ]; then
  for C in c1 c3 c4 c5; do
    timeout 900 python bench.py --config $C > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
    echo "bench $C exit $?"; cut -c1-300 $OUT/${TAG}_bench_$C.json; tail -2 $OUT/${TAG}_bench_$C.err
  done
  timeout 600 python bench.py --impl reference > $OUT/${TAG}_ref_c2.json 2>/dev/null
fi
SMALL="python bench.py --frames 256 --steps 1 --warmup 3 --no-cpu-baseline"
if $SMALL > $OUT/${TAG}_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'harris|boxsum|select|stereo_match|describe_left|bin_keypoints' -s 48 -c 96 \
      --csv --log-file $OUT/${TAG}_launches.csv $SMALL > $OUT/${TAG}_ncu_l.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harris|boxsum|select|stereo_match|describe_left|bin_keypoints' -s 48 -c 6 \
      -o $OUT/${TAG}_prof -f $SMALL > $OUT/${TAG}_ncu_f.log 2>&1
  echo "ncu full exit $?"
else
  echo "plain small run failed"; tail -20 $OUT/${TAG}_plain.log
fi
