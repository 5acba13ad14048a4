#!/bin/bash
# Developer tool: tests/build_variant.sh <name> "<extra nvcc flags>" -> build_variants/<name>.so (+ .ptxas.log), built from
# the working tree.  The variants travel to the GPU box with the snapshot; `tests/gpu_session.sh <tag> variants` times them.
set -eu
cd "$(dirname "$0")/.."
mkdir -p build_variants
make -s -C raytracegr.jl_b200/csrc rtgr_embedded.h tsit5_tables.h
cd raytracegr.jl_b200/csrc
nvcc $2 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2,-pthread --shared \
    -Xptxas -v -o ../../build_variants/$1.so raytracegr_cuda.cu -ldl 2> ../../build_variants/$1.ptxas.log
grep -A2 'trace_kernelILi1ELi0' ../../build_variants/$1.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' '
echo " <- $1 [$2]"
