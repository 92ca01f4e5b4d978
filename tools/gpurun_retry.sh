#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged):
#   tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$OUT"
  exit 0
done
echo "gave up: pod busy"
