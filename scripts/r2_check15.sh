#!/usr/bin/env bash
# round 2: last pass on the FINAL sources: tap-GEMM tests (incl. the ksub=1 variant), ncu launch list + tapgemm2h capture
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c15_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "tapgemm" > gpurun_out/r2_c15_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c15_tests.log | cut -c1-300 | tail -4 | tee -a $S
timeout 300 python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:"tc_tapgemm2h_kernel" -s 6 -c 4 -f -o gpurun_out/r2_prof_tapgemm2h python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_tapgemm2h.log 2>&1
echo "ncu tapgemm2h rc=$?" | tee -a $S
