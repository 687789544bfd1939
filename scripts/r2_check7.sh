#!/usr/bin/env bash
# round 2, seventh GPU pass (1 GPU): effect of sleeping epilogue waits; wgrad pair vs single per shape; config-2 bench both ways
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c7_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 300 python scripts/wgrad_bench.py 2>&1 | tee -a $S
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_ops_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "module_graphs or wgrad or tapgemm or clip or golden" > gpurun_out/r2_c7_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c7_tests.log | cut -c1-300 | tail -10 | tee -a $S
bench() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 900 python bench.py --steps 10 --warmup 3 --max-seconds 800 --no-cpu-baseline > gpurun_out/r2_c7_bench_$tag.json 2> gpurun_out/r2_c7_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c7_bench_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "loss")},
          "roofline", r.get("frac"), r.get("avg_launch_ms"), "wgrad", r.get("wgrad"), "e2e", d["e2e"]["ms_per_step"])
except Exception as e:
    print("no bench line:", e)
PY
}
bench single MEGB200_WGRAD_PAIR=0
bench pair MEGB200_WGRAD_PAIR=1
