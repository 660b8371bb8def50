#!/usr/bin/env bash
# Build libdd_b200.so for sm_100a (in-tree, next to the Python host layer).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/na-nonlinear-temperature-enhanced-diffusion-model-dd_b200/csrc"
OUT="${DD_OUT:-$HERE/na-nonlinear-temperature-enhanced-diffusion-model-dd_b200/libdd_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -Xcompiler -fvisibility=default --shared \
  ${DD_PTXAS_V:+-Xptxas -v} ${DD_EXTRA_FLAGS:-} \
  -o "$OUT" "$SRC/dd_kernels.cu" "$SRC/dd_solver.cu" "$SRC/dd_capi.cu" -lcudart
echo "built $OUT"
