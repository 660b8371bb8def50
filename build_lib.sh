#!/usr/bin/env bash
# Build libdd_b200.so for sm_100a (in-tree, next to the Python host layer).  The translation units are compiled
# in parallel into tests-free object files under build/ and linked into one shared library.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/na-nonlinear-temperature-enhanced-diffusion-model-dd_b200/csrc"
OUT="${DD_OUT:-$HERE/na-nonlinear-temperature-enhanced-diffusion-model-dd_b200/libdd_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="${DD_OBJ:-$HERE/build/obj}"
mkdir -p "$OBJ"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
       -Xcompiler -fvisibility=default ${DD_PTXAS_V:+-Xptxas -v} ${DD_EXTRA_FLAGS:-})
pids=()
for u in dd_kernels dd_solver dd_wave dd_lane dd_member dd_halo dd_capi; do
  "$NVCC" "${FLAGS[@]}" -c -o "$OBJ/$u.o" "$SRC/$u.cu" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" --shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/dd_kernels.o "$OBJ"/dd_solver.o \
  "$OBJ"/dd_wave.o "$OBJ"/dd_lane.o "$OBJ"/dd_member.o "$OBJ"/dd_halo.o "$OBJ"/dd_capi.o -lcudart
echo "built $OUT"
