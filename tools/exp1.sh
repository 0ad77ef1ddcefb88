#!/bin/bash
# round-2 session-3 experiments, one box: store-predicate A/B of the pyramid kernel, two-handle overlap probe, C4 compaction capture
out=gpurun_out/r2/exp1; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3 > $out/pytest.log
bash tools/pyr_ab2.sh $out 2 "" "_base" > /dev/null 2>&1
timeout 300 python tools/two_handles_probe.py 100 1:8 2:8 2:7 2:6 1:7 4:7 > $out/handles.log 2>&1
timeout 200 python tools/config_stages.py C4 > $out/c4_stages.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:compact_tracks -s 6 -c 1 -o $out/compact_c4 python tools/config_stages.py C4 12 > $out/ncu_compact.log 2>&1
cat $out/pytest.log $out/ab2.log $out/handles.log $out/c4_stages.log
