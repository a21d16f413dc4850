#!/usr/bin/env bash
# Build A/B variants of libpar2d_gpu.so with extra -D flags into par2d_b200/csrc/build/variants/.
# usage: tools/build_variants.sh name1 "-DFLAG=.." name2 "-D.." ...   then  PAR2D_GPU_LIB=<path> python bench.py
set -euo pipefail
cd "$(dirname "$0")/../par2d_b200/csrc"
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I../../include"
while [ $# -ge 2 ]; do
    name="$1"; flags="$2"; shift 2
    d="build/variants/$name"; mkdir -p "$d"
    $NV $flags -Xptxas -v -c kernels_fast.cu -o "$d/kernels_fast.o" 2> "$d/ptxas.txt" &
    $NV $flags -c kernels_march.cu -o "$d/kernels_march.o" &
    $NV $flags -fmad=false -c kernels_faithful.cu -o "$d/kernels_faithful.o" &
    $NV $flags -fmad=false -c kernels_common.cu -o "$d/kernels_common.o" &
    $NV $flags -c capi.cu -o "$d/capi.o" &
    wait
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "build/variants/libpar2d_gpu_$name.so" "$d"/*.o
    echo "$name: $(grep -A2 'march_kernelILb0ELb0ELb1' "$d/ptxas.txt" | grep -E 'registers' | head -1) / $(grep -A2 'march_kernelILb0ELb0ELb1' "$d/ptxas.txt" | grep -E 'spill' | head -1)"
done
