#!/bin/bash
# interleaved A/B of pyramid kernel builds (same box, order rotated): tools/pyr_ab2.sh <outdir> <reps> <variant suffixes...>
out=$1; reps=$2; shift 2; mkdir -p $out
for r in $(seq $reps); do
  for c in "C1 128" "C1 640" "C4 96" "EUROC6 128"; do
    for v in "$@"; do
      RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so timeout 300 python tools/pyr_bench.py $c 2>&1 | tail -1 | sed "s/^/[$v] /" >> $out/ab2.log
    done
  done
done
cat $out/ab2.log
