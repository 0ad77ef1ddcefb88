#!/bin/bash
# session-3 experiment 5: split temporal LK launch (bulk beside the previous frame's tail)
out=gpurun_out/r2/exp5; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/pytest.log
for r in 1 2; do
  RSVIO_NO_SPLIT_LK=1 timeout 300 python tools/two_handles_probe.py 100 1:8 2>&1 | tail -1 | sed "s/^/[nosplit] /" >> $out/ab.log
  timeout 300 python tools/two_handles_probe.py 100 1:8 2>&1 | tail -1 | sed "s/^/[split] /" >> $out/ab.log
done
cat $out/pytest.log $out/ab.log
