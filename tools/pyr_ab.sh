#!/bin/bash
# A/B of pyramid kernel builds on one GPU box: tools/pyr_ab.sh <outdir> <variant suffixes...> ("" = the default build)
out=${1:-gpurun_out/pyr_ab}; shift; mkdir -p $out
for v in "$@"; do
  export RSVIO_LIB=$PWD/rs-vio_b200/lib/librsvio_b200$v.so
  echo "== variant '$v'" >> $out/ab.log
  timeout 600 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_edge_cases.py tests/test_thirdparty_pins.py -m gpu -x -q 2>&1 | tail -3 >> $out/ab.log
  for c in "C1 128" "C1 640" "C2 128" "C3 128" "C4 24" "C4 96" "EUROC6 128" "EUROC6 2" "C1 2"; do
    timeout 300 python tools/pyr_bench.py $c 2>&1 | tail -2 >> $out/ab.log
  done
done
unset RSVIO_LIB
for c in "EUROC6 128" "C4 24" "EUROC6 2"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$out/launches_${c// /_}.csv" python tools/pyr_bench.py $c 3 > /dev/null 2>&1
done
cat $out/ab.log
