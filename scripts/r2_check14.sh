#!/usr/bin/env bash
# round 2: deeper ring (one k chunk per stage) in the plain pair kernel -- correctness, effect, and the ncu captures on these sources
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c14_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "tapgemm or config2 or golden" > gpurun_out/r2_c14_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c14_tests.log | cut -c1-300 | tail -8 | tee -a $S
bench() {
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --max-seconds 500 --no-e2e --no-cpu-baseline > gpurun_out/r2_c14_bench_$tag.json 2> gpurun_out/r2_c14_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c14_bench_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "loss", "step_frac_of_peak")}, "roofline", r.get("frac"), r.get("achieved"), r.get("avg_launch_ms"))
except Exception as e:
    print("no bench line:", e)
PY
}
bench auto MEGB200_TAPGEMM_KSUB=0
bench ksub2 MEGB200_TAPGEMM_KSUB=2
timeout 300 python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
cap() {
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r2_prof_$1 python bench.py --ncu-step --warmup 3 > gpurun_out/r2_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" | tee -a $S
}
cap tapgemm2h "tc_tapgemm2h_kernel" 6 4
cap tapgemm2 "tc_tapgemm2_kernel" 0 4
