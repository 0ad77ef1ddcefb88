#!/bin/bash
# session-3 experiment 6: helper launch in the split temporal LK
out=gpurun_out/r2/exp6; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3 > $out/pytest.log
for r in 1 2; do
  RSVIO_LK_HELPER=0 timeout 300 python tools/two_handles_probe.py 100 1:8 2>&1 | tail -1 | sed "s/^/[split] /" >> $out/ab.log
  timeout 300 python tools/two_handles_probe.py 100 1:8 2>&1 | tail -1 | sed "s/^/[split+helper] /" >> $out/ab.log
done
cat $out/pytest.log $out/ab.log
