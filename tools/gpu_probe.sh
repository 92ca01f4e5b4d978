#!/bin/bash
# Probes: parity tests, device-resident value of some configurations, stereo_match time against the pair table's distinct
# test points (three library builds), ncu launch list of tracking calls (C3).  usage: bash tools/gpu_probe.sh <tag>
TAG=${1:-probe}; OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for C in c1 c2 c5; do python bench.py --config $C --device-only 2>&1 | tail -1 | cut -c1-420; done
python tools/bench_tracking.py 2>&1 | tail -1 | cut -c1-300
for L in svi_mapper_b200/libsvi_gpu.so build/alt/alt_random/libsvi_gpu.so build/alt/alt_adversarial/libsvi_gpu.so; do
  SVI_GPU_LIB=$PWD/$L python bench.py --device-only --steps 3 2>/dev/null | tail -1 > $OUT/${TAG}_table_$(basename $(dirname $L)).json
  python -c "
import json,sys; d=json.loads(open('$OUT/${TAG}_table_$(basename $(dirname $L)).json').read()); print('$L', round(d['value']), d['stage_ms'])"
done
C3="python bench.py --config c3 --frames 9 --steps 1 --no-cpu-baseline"
if $C3 > $OUT/${TAG}_c3_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kernel' -s 250 -c 150 --csv --log-file $OUT/${TAG}_c3_launches.csv $C3 > $OUT/${TAG}_c3_ncu.log 2>&1
  echo "ncu c3 exit $?"
fi
