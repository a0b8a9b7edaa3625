#!/bin/sh
# Build a variant of the CUDA library for A/B runs:  tools/build_variant.sh NAME [-Dflags...]
# -> build/libnexo_NAME.so ; select it with NEXOCLOM_B200_LIB=build/libnexo_NAME.so
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$HERE/build"
cd "$HERE/nexoclom_b200/csrc"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC -shared "$@" nx_kernels.cu nx_los_grid.cu nx_source_map.cu nx_compact.cu nx_comm.cu nx_api.cu -ldl -o "$HERE/build/libnexo_$NAME.so"
echo "built build/libnexo_$NAME.so"
