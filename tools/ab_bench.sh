#!/bin/bash
# A/B two builds of libsvi_gpu.so on the same box: alternate runs, print value / e2e / stage shares.
# usage: tools/ab_bench.sh <libA> <libB> [rounds]
A=$1; B=$2; N=${3:-3}
for i in $(seq $N); do
  for L in $A $B; do
    SVI_GPU_LIB=$PWD/$L python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', round(d['value']), round(d['e2e']['value']), {k:round(v) for k,v in d['roofline']['stage_share'].items()})"
  done
done
