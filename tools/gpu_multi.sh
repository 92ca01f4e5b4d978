#!/bin/bash
# Multi-GPU session (8 GPUs of one box): in-library driver equivalence test, host->device bandwidth matrix, C2 weak scaling with
# and without NUMA binding / strided placement, C4 strong scaling.  usage: bash tools/gpu_multi.sh <tag>
set -u
TAG=${1:-multi}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu_driver" 2>&1 | tail -3
timeout 300 python tools/h2d_matrix.py > $OUT/${TAG}_h2d_matrix.json 2> $OUT/${TAG}_h2d.err; echo "h2d exit $?"
run() { # name nproc extra-args...
  local name=$1 n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $n "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "h2d GB/s per rank", round(d["e2e"]["h2d_gbs_achieved"]/d["n_gpus"],1), d["config"].get("cpu_affinity"))
except Exception as e:
    print("   no line:", e)
PY
}
python bench.py --steps 3 --no-cpu-baseline > $OUT/${TAG}_c2_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('$OUT/${TAG}_c2_n1.json').read().strip().splitlines()[-1]); print('c2_n1 value', round(d['value']), 'e2e', round(d['e2e']['value']))"
run c2_n2 2 --steps 3
run c2_n4 4 --steps 3
run c2_n8 8 --steps 3
run c2_n4_nobind 4 --steps 3 --no-numa-bind
run c2_n8_nobind 8 --steps 3 --no-numa-bind
run c2_n4_strided 4 --steps 3 --device-order 0,2,4,6
run c2_n2_far 2 --steps 3 --device-order 0,4
run c4_n2 2 --config c4 --steps 2
run c4_n4 4 --config c4 --steps 2
run c4_n8 8 --config c4 --steps 2
