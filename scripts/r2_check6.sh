#!/usr/bin/env bash
# round 2, sixth GPU pass (1 GPU): weight-gradient kernels per shape + ncu of both; the tests that failed in pass 5
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c6_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -k "module_graphs or wgrad or wide" > gpurun_out/r2_c6_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c6_tests.log | cut -c1-300 | tail -10 | tee -a $S
timeout 300 python scripts/wgrad_bench.py 2>&1 | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_wgrad" -f -o gpurun_out/r2_prof_wgrad \
    python scripts/wgrad_bench.py --ncu > gpurun_out/r2_c6_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $S
ls -la gpurun_out/r2_prof_wgrad.ncu-rep 2>&1 | tee -a $S
