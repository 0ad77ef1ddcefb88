#!/bin/bash
# The repo's stand-in for compute-sanitizer memcheck (closed on the GPU pool): the GPU parity tests against the
# -DRSVIO_BOUNDS_CHECK build, in which every unclamped gather / shared-memory tile index / scratch index of the
# pyramid and LK kernels is asserted (a violation prints its source line and traps).  Usage: tools/bounds_check.sh <outdir>
out=${1:-gpurun_out/bounds}; mkdir -p "$out"
lib=$PWD/rs-vio_b200/lib/librsvio_b200_check.so
[ -f "$lib" ] || { echo "build it first: make -C rs-vio_b200 VARIANT=check EXTRA=-DRSVIO_BOUNDS_CHECK"; exit 1; }
RSVIO_LIB=$lib timeout 1500 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_edge_cases.py tests/test_gpu_lk.py \
    tests/test_gpu_tracker.py tests/test_gpu_detect.py tests/test_thirdparty_pins.py -m gpu -q -p no:cacheprovider > "$out/pytest_check.log" 2>&1
echo "rc=$? violations=$(grep -c 'bounds check failed' "$out/pytest_check.log") $(tail -1 "$out/pytest_check.log")" | tee "$out/summary.txt"
