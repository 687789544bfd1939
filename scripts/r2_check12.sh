#!/usr/bin/env bash
# round 2: per-phase clock64 accounting of the pair tap-GEMM kernels after the elect.sync change (developer build)
set -u
mkdir -p gpurun_out
MEGB200_EXTRA_NVCC_FLAGS=-DMEGB200_DEBUG bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build_dev.log 2>&1 || { echo "DEV BUILD FAILED"; tail -5 gpurun_out/r2_build_dev.log; }
MEGB200_DEV_LIB=/tmp/libmegb200_dev.so MEGB200_DEBUG_TIMING=1 timeout 300 python scripts/step_profile.py gwilliams 256 > gpurun_out/r2_c12_step.log 2> gpurun_out/r2_c12_timing.log
echo "rc=$?"
grep "megb200 timing" gpurun_out/r2_c12_timing.log | sort | uniq -c | sort -rn | head -40 | cut -c1-420
