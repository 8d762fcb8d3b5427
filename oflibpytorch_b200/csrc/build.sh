#!/bin/sh
# Builds liboflib_b200.so in-tree for sm_100a.  -fmad=false: every FMA in the kernels is explicit
# (see common.cuh); -lineinfo so ncu's source page maps to these files.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC -shared ${OFB_NVCC_EXTRA} \
    -o liboflib_b200.so api.cu stats.cu warp_t.cu warp_t_adj.cu warp_t_tile.cu warp_t_inset.cu splat_s.cu track.cu ingest.cu rowband.cu
