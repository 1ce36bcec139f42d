#!/usr/bin/env bash
# Builds Oracle 1: the UNMODIFIED reference renderer core, compiled from the sources where they
# lie under $GVV_REFERENCE (default /root/reference), plus oracle/ref_harness.cu.  Output only
# into oracle/_ref/ (git-ignored, shipped to the GPU box).  The reference's own build system
# (cmake + TensorFlow) is not used; TensorFlow is not installable offline.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GVV_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/cpp/src/Renderer" ]; then echo "reference not found at $REF" >&2; exit 3; fi
mkdir -p "$OUT"
SRC="$REF/cpp/src"
INC="-I$SRC -I$REF/cpp/thirdParty/Shared/cutil/inc"
# same effective flags as the reference's CMake (only -arch reaches nvcc: cmakeTF2Linux/CMakeLists.txt:107-108),
# retargeted to sm_100a
FLAGS="-O3 -gencode arch=compute_100a,code=sm_100a -std=c++17 -w -Xcompiler -fPIC,-fopenmp $INC"
nvcc $FLAGS -c "$SRC/Renderer/CUDABasedRasterization.cu" -o "$OUT/ras.o"
nvcc $FLAGS -c "$SRC/Renderer/CUDABasedRasterizationGrad.cu" -o "$OUT/rasgrad.o"
nvcc $FLAGS -x cu -c "$SRC/Renderer/CUDABasedRasterization.cpp" -o "$OUT/ras_host.o"
nvcc $FLAGS -x cu -c "$SRC/Renderer/CUDABasedRasterizationGrad.cpp" -o "$OUT/rasgrad_host.o"
nvcc $FLAGS -c "$HERE/ref_harness.cu" -o "$OUT/harness.o"
nvcc -shared -o "$OUT/libgvv_ref.so" "$OUT/ras.o" "$OUT/rasgrad.o" "$OUT/ras_host.o" "$OUT/rasgrad_host.o" "$OUT/harness.o" -Xcompiler -fopenmp -lgomp
rm -f "$OUT"/*.o
echo "built $OUT/libgvv_ref.so"
