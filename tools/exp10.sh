#!/bin/bash
# session-3 experiment 10: 32- and 80-row tiles of the pyramid tile kernel (single-buffered) against the 48-row default
out=gpurun_out/r2/exp10; mkdir -p $out
RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200_th32.so timeout 300 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -1 > $out/pytest_th32.log
for c in "C1 128" "C1 640" "EUROC6 128" "C4 96"; do
  for v in "" "_th32" "_th80"; do
    RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so timeout 120 python tools/pyr_bench.py $c 2>&1 | tail -1 | sed "s/^/[$v] /" >> $out/ab.log
  done
done
cat $out/pytest_th32.log $out/ab.log
