#!/bin/bash
# latency-path session: parity tests (optionally a -k filter), single-pair / tracking benches
timeout 1500 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -25
python bench.py --config c1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c1', d['value'], d['e2e'])"
python bench.py --config c3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['config']['ms_per_tracking_frame_median'], d['config']['ms_per_detection_frame_median'])"
python tools/bench_latency.py 2>&1 | tail -1
python tools/bench_tracking.py 2>&1 | tail -2
SVI_TRACE=1 python tools/bench_tracking.py 2>&1 | grep "svi_track_landmarks n=" | tail -4
