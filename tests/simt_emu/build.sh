#!/bin/bash
# usage: build.sh <output dir>   -- host build of the RDM kernels under the SIMT emulator (test infrastructure)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${1:-$HERE/_build}
mkdir -p "$OUT"
# the kernel text, with the two shared-memory declaration forms mapped to host storage and the CUDA include dropped
sed -e 's/^\s*extern __shared__ double rdm_smem\[\];//' \
    -e 's/__shared__ /static /' \
    -e 's/#include <cuda_runtime.h>//' \
    "$HERE/../../forte2_b200/csrc/rdm_kernels.cuh" > "$OUT/rdm_kernels_host.cuh"
${CXX:-g++} -std=c++20 -O1 -fPIC -shared -pthread -I"$HERE" -I"$OUT" "$HERE/rdm_emu.cc" -o "$OUT/librdm_emu.so" \
    -L/usr/lib/gcc/x86_64-linux-gnu/13
echo "$OUT/librdm_emu.so"
