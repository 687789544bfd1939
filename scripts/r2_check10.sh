#!/usr/bin/env bash
# round 2: side-stream overlap of the small fp32 launches -- correctness (graph / parity / module tests) and the bench both ways
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c10_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_parity_gpu.py tests/test_optim_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_c10_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c10_tests.log | cut -c1-300 | tail -8 | tee -a $S
bench() {
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --max-seconds 500 --no-e2e > gpurun_out/r2_c10_bench_$tag.json 2> gpurun_out/r2_c10_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c10_bench_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "loss")}, "traffic", r.get("traffic"), r.get("ncu_tensor_pipe_active_pct"), "e2e", d["e2e"]["ms_per_step"], "parity", (d.get("parity") or {}).get("z_rel"))
except Exception as e:
    print("no bench line:", e)
PY
}
bench side MEGB200_SIDE_STREAM=1
bench noside MEGB200_SIDE_STREAM=0
