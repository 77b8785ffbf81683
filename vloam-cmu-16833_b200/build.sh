#!/bin/bash
# Builds libvloam_b200.so (the C-ABI library, include/vloam_b200.h) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
OBJS=""
for f in sr_kernels lo_kernels lm_kernels vo_kernels capi; do
  if [ ! -f lib/$f.o ] || [ csrc/$f.cu -nt lib/$f.o ] || [ csrc/common.cuh -nt lib/$f.o ] || [ csrc/internal.h -nt lib/$f.o ] || [ ../include/vloam_b200.h -nt lib/$f.o ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS -c csrc/$f.cu -o lib/$f.o 2> lib/$f.ptxas.log || { cat lib/$f.ptxas.log; exit 1; }
  fi
  OBJS="$OBJS lib/$f.o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o lib/libvloam_b200.so $OBJS
echo "built lib/libvloam_b200.so"
