#!/bin/bash
# Build an experimental variant of the trace kernels: scripts/build_variant.sh NAME [-DVP_... flags]
# -> build_var/libvp_NAME.so (load it with VOLPRIM_CUDA_LIB=$PWD/build_var/libvp_NAME.so)
set -e
cd "$(dirname "$0")/../volprim_balance_b200/csrc"
name=$1; shift
mkdir -p ../../build_var
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I. \
     --expt-relaxed-constexpr "$@" -Xptxas -v -c vp_trace.cu -o /tmp/vt_$name.o 2>/tmp/vt_$name.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_var/libvp_$name.so _build/vp_api.o _build/vp_build.o \
     _build/vp_optim.o _build/vp_film.o /tmp/vt_$name.o -lcudart_static -ldl -lrt -lpthread
