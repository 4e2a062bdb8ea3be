#!/bin/bash
# Build a variant of the library with extra -D flags on the kernel sources (experiments):
#   tools/build_variant.sh stamps -DPC_STAMPS_HEADER='"../../tools/debug/stamps.cuh"'
#       ->  variants/libpunch_stamps.so   (use with PUNCH_B200_LIB=...; phase stamps of tools/debug/stamps.cuh,
#           which the product library never includes)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude"
nvcc $F "$@" -c clockpunch_b200/csrc/kernels_ukf.cu -o variants/ukf_$name.o &
nvcc $F "$@" -c clockpunch_b200/csrc/kernels_propagate.cu -o variants/prop_$name.o &
nvcc $F "$@" -c clockpunch_b200/csrc/kernels_env.cu -o variants/env_$name.o &
nvcc $F "$@" -c clockpunch_b200/csrc/api.cu -o variants/api_$name.o &
wait
nvcc -shared -o variants/libpunch_$name.so variants/ukf_$name.o variants/prop_$name.o variants/env_$name.o \
  clockpunch_b200/build/kernels_vis.o variants/api_$name.o \
  -gencode arch=compute_100a,code=sm_100a -cudart static
echo variants/libpunch_$name.so
