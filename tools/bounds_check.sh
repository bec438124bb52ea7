#!/bin/bash
# compute-sanitizer is not available on the GPU pool.  This builds the whole library with -DESSA_BOUNDS (device-side assertions
# on computed buffer addresses and indices: ESSA_CHECK in common.cuh -- the slots / partials / keys / permutation of the
# pair-shared kernels, the shared-memory indices of the grid kernel, every History and Trail ring write of the recorders that
# all resident kernels share, the snapshot ring and the links of the 2-SM kernel; a failed one prints file:line and traps) into
# build/variants/lib_bounds.so and runs the GPU tests under it.
# usage:  tools/bounds_check.sh [--build-only]     (build here or on the GPU box; run on a GPU box)
set -e
cd "$(dirname "$0")/.."
python -m essa_b200.build > /dev/null
mkdir -p build/variants
pids=""
for f in essa_b200/csrc/*.cu; do
    src=$(basename $f)
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --fmad=true -DESSA_BOUNDS \
        -c $f -o build/variants/${src}_bounds.o &
    pids="$pids $!"
done
for p in $pids; do wait $p; done
rest=$(ls essa_b200/lib/obj/*.cpp.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_bounds.so $rest build/variants/*_bounds.o -ldl -cudart static
echo "built build/variants/lib_bounds.so"
[ "$1" = "--build-only" ] && exit 0
ESSA_B200_LIB=build/variants/lib_bounds.so python -m pytest tests -x -q -m gpu
