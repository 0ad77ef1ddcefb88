#!/bin/bash
# session-3 experiment 8: split threshold, narrow detection grid, split off while upload-bound
out=gpurun_out/r2/exp8; mkdir -p $out
export RSVIO_FRAME_CACHE=/tmp/rsvio_frames
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_detect.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3 > $out/pytest.log
B="python bench.py --steps 100 --no-cpu-baseline --no-pyramid-roofline --no-per-config"
for S in 32 64; do
  for ns in 1 0; do
    RSVIO_NO_SPLIT_LK=$ns timeout 300 $B --streams-per-gpu $S 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('S=$S nosplit=$ns', 'ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,1), 'stages', {k: round(v*1e3,1) for k,v in d['stage_ms_per_step'].items()})" >> $out/ab.log
  done
done
cat $out/pytest.log $out/ab.log
