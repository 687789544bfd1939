#!/usr/bin/env bash
# round 2, final evidence pass (1 GPU) on the final sources: full GPU suite, smoke, bench (default flags), ncu launch list +
# full captures of the dominant kernels
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_final2_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_final2_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_final2_tests.log | cut -c1-300 | tail -8 | tee -a $S
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_final2_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -1 gpurun_out/r2_final2_smoke.log | cut -c1-400 | tee -a $S
timeout 900 python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err
echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_final2_bench_ref.json 2> gpurun_out/r2_final2_bench_ref.err
echo "bench reference rc=$?" | tee -a $S
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
cap wgrad "tc_wgrad_kernel" 2 3
cap clip "tc_clip" 0 2
python - <<PY | tee -a $S
import json
d = json.loads(open("gpurun_out/r2_final2_bench.json").read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "loss", "step_frac_of_peak", "gpu_launches", "clocks")})
print("roofline", {k: r.get(k) for k in ("achieved", "frac", "avg_launch_ms", "share_of_step", "traffic", "wgrad")})
print("e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "h2d_bytes_per_step", "h2d_alone_ms")})
print("parity", d.get("parity"))
print("cpu", d.get("cpu_baseline"))
print("ref", open("gpurun_out/r2_final2_bench_ref.json").read()[:300])
PY
