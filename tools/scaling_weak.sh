#!/bin/bash
# One weak-scaling point (64 streams per GPU) on the N GPUs of this box: tools/scaling_weak.sh <outdir> <N> [bench args]
out=$1; n=$2; shift 2; mkdir -p "$out"
name="weak_n$n$( [ -n "$1" ] && echo "_$(echo "$*" | tr -cd 'a-z0-9')" )"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --steps 80 --warmup 10 --no-per-config "$@" > "$out/$name.json" 2> "$out/$name.err"
echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('$out/$name.json').read().strip().splitlines()[-1])
    print('value %.1fM e2e %.1fM ms %.3f e2e_ms %.3f h2d_all %.0f GB/s pyr_frac %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_all_gpus'], d['roofline']['frac']))
except Exception as e: print('no line', e)
")" | tee -a "$out/summary.txt"
