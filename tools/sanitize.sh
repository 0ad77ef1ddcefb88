#!/bin/bash
# compute-sanitizer passes over the aliasing / unclamped-gather code: memcheck, racecheck, synccheck.
# Usage (on a GPU box): tools/sanitize.sh <outdir>.  Each summary ends with the tool's ERROR SUMMARY line.
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
PYR='tests/test_gpu_pyramid.py -k "C1 or odd or deep or EUROC or golden"'
LK='tests/test_gpu_lk.py::test_lk_random_points_including_failures'
SEQ='tests/test_gpu_tracker.py::test_remove_id_and_reset tests/test_gpu_detect.py::test_detect_on_noise_and_flat'
for tool in memcheck racecheck synccheck; do
  for name in PYR LK SEQ; do
    sel=${!name}
    eval timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest $sel -x -q -p no:cacheprovider > "$out/${tool}_${name}.log" 2>&1
    echo "$tool $name rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$out/${tool}_${name}.log" | tr '\n' ' ')" >> "$out/summary.txt"
  done
done
cat "$out/summary.txt"
