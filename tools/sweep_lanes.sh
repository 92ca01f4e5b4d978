#!/bin/bash
# Sweep lanes x chunk size of the stereo pipeline on one GPU (device-resident value and host e2e).
for L in ${LANES:-4 6 8}; do for C in ${CHUNKS:-16 32 64}; do
  SVI_LANES=$L SVI_CHUNK_FRAMES=$C python bench.py --no-cpu-baseline --steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lanes $L chunk $C value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
