#!/usr/bin/env bash
# round 2: L2 prefetch of the next tile's activations in the pair tap-GEMM producers -- correctness + effect
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c13_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "tapgemm or config2 or golden" > gpurun_out/r2_c13_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c13_tests.log | cut -c1-300 | tail -8 | tee -a $S
bench() {
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --max-seconds 500 --no-e2e --no-cpu-baseline > gpurun_out/r2_c13_bench_$tag.json 2> gpurun_out/r2_c13_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c13_bench_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "loss", "step_frac_of_peak")}, "roofline", r.get("frac"), r.get("achieved"), r.get("avg_launch_ms"))
except Exception as e:
    print("no bench line:", e)
PY
}
bench prefetch MEGB200_TAPGEMM_L2_PREFETCH=1
bench noprefetch MEGB200_TAPGEMM_L2_PREFETCH=0
