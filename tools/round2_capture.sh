#!/bin/bash
# Round-2 evidence on one GPU box: full GPU test suite, bounds-check build, soak, the default bench line, ncu launch
# list + full captures, per-kernel lists of the pyramid shapes and of a crate B frame.  Outputs under gpurun_out/r2/final.
o=${1:-gpurun_out/r2/final}; mkdir -p $o
export RSVIO_FRAME_CACHE=/tmp/rsvio_frames
python -m pytest tests -m gpu -q > $o/pytest_gpu.log 2>&1; tail -2 $o/pytest_gpu.log
cp gpurun_out/parity_stats.jsonl $o/parity_stats.jsonl 2>/dev/null
bash tools/bounds_check.sh $o/bounds
timeout 600 python tools/soak.py 2000 > $o/soak.log 2>&1; tail -2 $o/soak.log
timeout 300 python tools/crate_b_probe.py 300 > $o/crate_b_probe.log 2>&1; tail -1 $o/crate_b_probe.log
env -u RSVIO_FRAME_CACHE python bench.py > $o/bench.json 2> $o/bench.err; tail -c 300 $o/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_reference.json 2> $o/bench_reference.err
bash tools/ncu_capture.sh $o/ncu > $o/ncu_capture.log 2>&1
for c in "EUROC6 128" "C4 24" "EUROC6 2" "C1 128"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$o/launches_${c// /_}.csv" python tools/pyr_bench.py $c 3 > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_crate_b.csv python tools/crate_b_probe.py 8 > /dev/null 2>&1
ls $o $o/ncu
