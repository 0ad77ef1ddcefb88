#!/bin/bash
# session-3 experiment 2: chunked compaction kernel + exact vertical tents in the generic pyramid kernel
out=gpurun_out/r2/exp2; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/pytest.log
for r in 1 2; do
  for c in "EUROC6 128" "EUROC6 2" "C4 96" "C3 128"; do
    for v in "" "_base"; do
      RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so timeout 300 python tools/pyr_bench.py $c 2>&1 | tail -1 | sed "s/^/[$v] /" >> $out/ab.log
    done
  done
done
for c in C4 C1 C3 EUROC_SHIPPED; do timeout 200 python tools/config_stages.py $c >> $out/stages.log 2>&1; done
cat $out/pytest.log $out/ab.log $out/stages.log
