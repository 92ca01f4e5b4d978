#!/bin/bash
# One GPU-box session: parity tests, the default bench line, the ncu launch list and one --set full capture of the
# four kernels of the new-landmark path.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
cat $OUT/${TAG}_bench.json
SMALL="python bench.py --frames 256 --steps 1 --warmup 3 --no-cpu-baseline"
if $SMALL > $OUT/${TAG}_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'harris|boxsum|select|stereo_match' -s 32 -c 64 \
      --csv --log-file $OUT/${TAG}_launches.csv $SMALL > $OUT/${TAG}_ncu_l.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harris|boxsum|select|stereo_match' -s 32 -c 4 \
      -o $OUT/${TAG}_prof -f $SMALL > $OUT/${TAG}_ncu_f.log 2>&1
  echo "ncu full exit $?"
else
  echo "plain small run failed"; tail -20 $OUT/${TAG}_plain.log
fi
