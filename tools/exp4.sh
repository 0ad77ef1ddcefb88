#!/bin/bash
# session-3 experiment 4: LK without XU-pipe conversions (I2FP for the pixels, FADD.RM floor), interleaved A/B on one box
out=gpurun_out/r2/exp4; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_lk.py tests/test_gpu_tracker.py -m gpu -x -q 2>&1 | tail -4 > $out/pytest.log
export RSVIO_PROBE_DEVICE_ONLY=1
for r in 1 2 3; do
  for v in "" "_nofloor" "_lkbase"; do
    RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so timeout 300 python tools/stage_probe.py 64 60 2>&1 | tail -1 | sed "s/^/[$v] /" >> $out/ab.log
  done
done
timeout 200 python tools/config_stages.py C4 >> $out/stages.log 2>&1
cat $out/pytest.log $out/ab.log $out/stages.log
