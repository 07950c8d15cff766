#!/bin/bash
# dev/build_variant.sh <name> <extra nvcc flags...>  -> cpu-path-tracing_b200/build/libptb200_<name>.so
set -e
cd "$(dirname "$0")/../cpu-path-tracing_b200"
name=$1; shift
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
mkdir -p build
$NV "$@" -c csrc/ptb_f32.cu -o build/ptb_f32_$name.o
$NV -shared -o build/libptb200_$name.so build/ptb_f32_$name.o build/ptb_f64.o build/ptb_smallpt_f64.o build/ptb_resolve.o build/ptb_api.o build/ptb_multi.o build/ptb_jit.o build/ptb_host.o -cudart static -ldl -lpthread
echo built build/libptb200_$name.so
