#!/bin/bash
# Multi-GPU runs on ONE box (gpurun --gpus 8): the strong-scaling curve BASELINE configs[4] names (64 streams in total,
# block-partitioned over N ranks) at N = 8, 4, 2 and the weak-scaling point at N = 8.  One JSON line per run in <outdir>.
out=${1:-gpurun_out/scaling}; mkdir -p "$out"
run() {  # name N extra-args...
  name=$1; n=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $n --steps 80 --warmup 10 --no-per-config "$@" > "$out/$name.json" 2> "$out/$name.err"
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$out/$name.json').read().strip().splitlines()[-1])
    print('value %.1fM e2e %.1fM ms %.3f e2e_ms %.3f h2d_all %.0f GB/s pyr_frac %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_all_gpus'], d['roofline']['frac']))
except Exception as e: print('no line', e)
")" | tee -a "$out/summary.txt"
}
run strong_n8 8 --total-streams 64
run strong_n4 4 --total-streams 64
run strong_n2 2 --total-streams 64
run weak_n8 8
[ "$2" = "pinned" ] && run weak_n8_pinned 8 --host-memory pinned
true
