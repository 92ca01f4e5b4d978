#!/bin/bash
# Harris tile width A/B: full parity suite on the default build, then device-resident values of both builds
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/ab_value.sh 2 svi_mapper_b200/libsvi_gpu.so build/libsvi_gpu_t64.so
for L in svi_mapper_b200/libsvi_gpu.so build/libsvi_gpu_t64.so; do
  for C in c1 c4 c5; do SVI_GPU_LIB=$PWD/$L python bench.py --config $C --device-only 2>&1 | tail -1 | cut -c1-330; done
done
