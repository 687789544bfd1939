#!/usr/bin/env bash
# Builds libmegb200.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../megb200/libmegb200.so"
obj="${here}/../../build/obj"
mkdir -p "${obj}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -Xptxas -v ${MEGB200_EXTRA_NVCC_FLAGS:-})
# developer builds (e.g. MEGB200_EXTRA_NVCC_FLAGS=-DMEGB200_DEBUG for the per-phase clock64 accounting) go to their own
# object directory / library so that they can never be mistaken for the product library
if [[ -n "${MEGB200_EXTRA_NVCC_FLAGS:-}" ]]; then
  obj="${obj}_dev"; out="${MEGB200_DEV_OUT:-/tmp/libmegb200_dev.so}"; mkdir -p "${obj}"
fi
pids=()
for f in api elementwise simt clip tc_gemm; do
  src="${here}/${f}.cu"
  o="${obj}/${f}.o"
  if [[ ! -f "${o}" || "${src}" -nt "${o}" || "${here}/common.cuh" -nt "${o}" || "${here}/tc_common.cuh" -nt "${o}" \
        || "${here}/../../include/megb200.h" -nt "${o}" ]]; then
    ( "${NVCC}" "${FLAGS[@]}" -c "${src}" -o "${o}" > "${obj}/${f}.log" 2>&1 || { cat "${obj}/${f}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -shared -gencode arch=compute_100a,code=sm_100a -o "${out}" "${obj}"/{api,elementwise,simt,clip,tc_gemm}.o -lcudart_static -ldl -lpthread -lrt
echo "built ${out}"
