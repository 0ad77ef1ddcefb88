#!/bin/bash
o=gpurun_out/r2/pyr5; mkdir -p $o
for r in 1 2 3; do
 for c in "C1 128" "C1 640" "C4 96"; do
  for v in "" _th32 _th24; do
   for nb in 1 2; do
     RSVIO_PYR_NBUF=$nb RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so timeout 300 python tools/pyr_bench.py $c 2>&1 | tail -1 | sed "s/^/[$v nbuf=$nb] /" >> $o/nbuf.log
   done
  done
 done
done
sort $o/nbuf.log
