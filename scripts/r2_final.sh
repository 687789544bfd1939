#!/usr/bin/env bash
# round 2, evidence pass (1 GPU): full GPU suite, smoke, compute-sanitizer on the tcgen05 kernels (small shapes), ncu launch
# list of one config-2 step + `ncu --set full` captures of the dominant kernels, final bench line.
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_final_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_final_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_final_tests.log | cut -c1-300 | tail -12 | tee -a $S
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -2 gpurun_out/r2_final_smoke.log | cut -c1-400 | tee -a $S
# ---- bench (before anything runs under a profiler)
timeout 900 python bench.py --steps 20 --warmup 5 --max-seconds 800 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?" | tee -a $S
tail -c 2500 gpurun_out/r2_final_bench.json | tee -a $S
# ---- compute-sanitizer: memcheck over every tensor-core kernel form on small shapes; racecheck on the CTA-pair tap-GEMM
SAN="tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -x"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest $SAN \
   -k "(forced_variants and case1) or (test_wgrad and case0 and tc) or (cta_pair_kernel_forced and case0) or (clip_pipeline and case0 and tc) or (fused_never and case0) or (two_row_tiles and case0)" \
   > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a $S; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2_sanitizer_memcheck.log | tail -5 | cut -c1-200 | tee -a $S
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest $SAN \
   -k "(forced_variants and case1 and (pair_halo or pair_plain or single-))" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a $S; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_sanitizer_racecheck.log | tail -5 | cut -c1-200 | tee -a $S
# ---- ncu: launch list of exactly one step, then full captures of the dominant kernels (each only after the plain run passed)
timeout 300 python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
cap() {  # tag, regex, skip, count
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r2_prof_$1 python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" | tee -a $S
}
cap tapgemm2h "tc_tapgemm2h_kernel" 6 4
cap tapgemm2 "tc_tapgemm2_kernel" 0 4
cap clip "tc_clip" 0 2
cap rowstage "rowstage_kernel" 4 6
ls -la gpurun_out/*.ncu-rep 2>&1 | tee -a $S
