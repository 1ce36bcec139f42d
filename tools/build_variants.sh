#!/bin/bash
# A/B builds of the library with compile-time switches: tools/build_variants.sh NAME "-DFLAG=0 ..." [NAME FLAGS ...]
# -> gpurun_out/variants/libgvv_NAME.so (travels to the GPU box with the snapshot? no: gpurun_out is not sent -> use gvv_variants/)
set -e
cd "$(dirname "$0")/.."
mkdir -p gvv_variants
S=gvv_differentiable_cuda_renderer_b200/csrc
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $2 -o gvv_variants/libgvv_$1.so \
    $S/gvv_api.cu $S/gvv_forward.cu $S/gvv_backward.cu $S/gvv_normalmap.cu $S/gvv_helpers.cu $S/gvv_microbench.cu &
  shift 2
done
wait
ls -la gvv_variants
