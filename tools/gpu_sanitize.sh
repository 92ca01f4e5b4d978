#!/bin/bash
# compute-sanitizer over three parity tests (one tool per GPU session, as the profiling guide asks):
#   bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck
TOOL=$1
OUT=gpurun_out
mkdir -p $OUT
K="test_stereo_frame_parity or test_track_manual_stage2_window_search or test_stress_frame_global_select_and_long_scanlines"
timeout 1500 compute-sanitizer --tool $TOOL --log-file $OUT/sanitizer_$TOOL.log python -m pytest tests/test_gpu_parity.py -x -q -k "$K" > $OUT/sanitizer_${TOOL}_pytest.log 2>&1
echo "exit $?"; tail -3 $OUT/sanitizer_${TOOL}_pytest.log; tail -5 $OUT/sanitizer_$TOOL.log; grep -c "=========" $OUT/sanitizer_$TOOL.log
