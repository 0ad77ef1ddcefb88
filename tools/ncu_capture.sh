#!/bin/bash
# One ncu call (GPU box): launch list of a short bench run + full captures (with source) of the pyramid kernel and of two consecutive LK launches
# (the launches of one frame; tools/ncu_summary.py reports the longest, the temporal bulk).
# Usage: tools/ncu_capture.sh <outdir> [bench args].  Numbers printed by runs under ncu are never bench values.
out=${1:-gpurun_out/ncu}; shift
mkdir -p "$out"
export RSVIO_FRAME_CACHE=${RSVIO_FRAME_CACHE:-/tmp/rsvio_frames}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pyramid-roofline --no-per-config $*"
$CMD > "$out/plain.log" 2>&1 || { echo "plain run failed"; tail -5 "$out/plain.log"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" $CMD > "$out/ncu_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:pyr_fast -s 4 -c 1 -o "$out/pyr" $CMD > "$out/ncu_pyr.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:lk_track -s 8 -c 4 -o "$out/lk" $CMD > "$out/ncu_lk.log" 2>&1
ls -la "$out"
