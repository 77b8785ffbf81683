#!/bin/bash
# Builds libvloam_b200.so (the C-ABI library, include/vloam_b200.h) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
mkdir -p lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
OBJS=""
HDRS="csrc/common.cuh csrc/internal.h csrc/gn_solver.cuh csrc/cta_sort.cuh csrc/gn_split.cuh csrc/gn_jet.cuh csrc/tma_bulk.cuh csrc/fdlibm_atan2f.h csrc/orb_pattern.inc ../include/vloam_b200.h"
for f in sr_kernels lo_kernels lm_kernels vo_kernels vo_detect vo_orb gn_split capi; do
  stale=0
  [ -f lib/$f.o ] || stale=1
  for d in csrc/$f.cu $HDRS; do [ "$d" -nt lib/$f.o ] && stale=1; done
  if [ $stale = 1 ]; then
    echo "nvcc $f.cu"
    $NVCC $FLAGS -c csrc/$f.cu -o lib/$f.o 2> lib/$f.ptxas.log || { cat lib/$f.ptxas.log; exit 1; }
  fi
  OBJS="$OBJS lib/$f.o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o lib/libvloam_b200.so $OBJS
echo "built lib/libvloam_b200.so"
