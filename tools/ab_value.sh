#!/bin/bash
# A/B builds of libsvi_gpu.so on the same box, device-resident value only: tools/ab_value.sh <rounds> <libA> <libB> ...
N=$1; shift
for i in $(seq $N); do
  for L in "$@"; do
    SVI_GPU_LIB=$PWD/$L python bench.py --device-only --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', round(d['value']), {k:round(v,2) for k,v in d['stage_ms'].items()})"
  done
done
