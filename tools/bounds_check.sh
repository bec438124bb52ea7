#!/bin/bash
# compute-sanitizer is not available on the GPU pool.  This builds the library with -DESSA_BOUNDS (device-side assertions on
# every computed buffer address / index of the pair-shared and grid kernels: ESSA_CHECK in common.cuh; a failed one prints
# file:line and traps) and runs those kernels' GPU tests under it.  usage (on a GPU box):  tools/bounds_check.sh
set -e
cd "$(dirname "$0")/.."
python -m essa_b200.build > /dev/null
mkdir -p build/variants
objs=""
for src in force_paired.cu force_grid.cu; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --fmad=true -DESSA_BOUNDS \
        -c essa_b200/csrc/$src -o build/variants/${src}_bounds.o
    objs="$objs build/variants/${src}_bounds.o"
done
rest=$(ls essa_b200/lib/obj/*.o | grep -v "/force_paired.cu.o" | grep -v "/force_grid.cu.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_bounds.so $rest $objs -ldl -cudart static
echo "built build/variants/lib_bounds.so"
[ "$1" = "--build-only" ] && exit 0
ESSA_B200_LIB=build/variants/lib_bounds.so python -m pytest tests/test_gpu_paired.py tests/test_gpu_grid.py tests/test_gpu_fullsize.py -x -q -m gpu
