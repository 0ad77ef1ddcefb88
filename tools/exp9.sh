#!/bin/bash
# session-3 experiment 9: the new split-launch test and a device-resident soak of the split pipeline
out=gpurun_out/r2/exp9; mkdir -p $out
export RSVIO_FRAME_CACHE=/tmp/rsvio_frames
timeout 600 python -m pytest tests/test_gpu_tracker.py -m gpu -x -q -k "split or c5_full or pipelined" 2>&1 | tail -3 > $out/pytest.log
timeout 600 python tools/soak.py 3000 32 device > $out/soak_split.log 2>&1
cat $out/pytest.log; tail -4 $out/soak_split.log
