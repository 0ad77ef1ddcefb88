#!/bin/bash
o=gpurun_out/r2/c1; mkdir -p $o
bash tools/bounds_check.sh $o/bounds
python bench.py > $o/bench.json 2> $o/bench.err; tail -c 600 $o/bench.err
bash tools/ncu_capture.sh $o/ncu > $o/ncu_capture.log 2>&1
for c in "EUROC6 128" "C4 24" "EUROC6 2" "C1 128"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$o/launches_${c// /_}.csv" python tools/pyr_bench.py $c 3 > /dev/null 2>&1
done
ls $o $o/ncu
