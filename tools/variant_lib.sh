#!/bin/bash
# Builds build/variants/lib_<name>.so: the library with ONE source recompiled under extra macros (kernel experiments; select
# with ESSA_B200_LIB=...).  usage: tools/variant_lib.sh <name> <source.cu> [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p build/variants
python -m essa_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --fmad=true "$@" \
    -c essa_b200/csrc/$src -o build/variants/${src}_$name.o
objs=$(ls essa_b200/lib/obj/*.o | grep -v "/$src.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/${src}_$name.o -ldl -cudart static
echo build/variants/lib_$name.so
