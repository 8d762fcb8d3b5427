#!/bin/sh
# Builds liboflib_b200.so in-tree for sm_100a.  -fmad=false: every FMA in the kernels is explicit
# (see common.cuh); -lineinfo so ncu's source page maps to these files.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC -shared ${OFB_NVCC_EXTRA} \
    -o liboflib_b200.so api.cu stats.cu warp_t.cu warp_t_adj.cu warp_t_tile.cu warp_t_inset.cu splat_s.cu track.cu ingest.cu rowband.cu
# The TORCH_LIBRARY shim for the two hot W ops (torch_binding.cpp): C++, no CUDA code of its own; links the library above.
TI=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))" 2>/dev/null | tail -1)
${CXX:-g++} -O2 -std=c++17 -shared -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 torch_binding.cpp \
    -I"$TI/include" -I"$TI/include/torch/csrc/api/include" -I/usr/local/cuda/include \
    -L"$TI/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -L. -loflib_b200 -Wl,-rpath,'$ORIGIN' \
    -o liboflib_b200_torch.so
