#!/usr/bin/env bash
# round 2: elect.sync single-thread roles (no per-instruction uniformisation loops) -- correctness + effect
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c11_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py tests/test_graph_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_c11_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier" gpurun_out/r2_c11_tests.log | cut -c1-300 | tail -8 | tee -a $S
timeout 300 python scripts/wgrad_bench.py 2>&1 | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 --max-seconds 500 --no-e2e > gpurun_out/r2_c11_bench.json 2> gpurun_out/r2_c11_bench.err
echo "bench rc=$?" | tee -a $S
python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c11_bench.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "loss", "step_frac_of_peak")}, "roofline", r.get("frac"), r.get("achieved"), r.get("avg_launch_ms"), "wgrad", r.get("wgrad"), "e2e", d["e2e"]["ms_per_step"], "parity", (d.get("parity") or {}).get("z_rel"))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/r2_c11_bench.err | cut -c1-300 | tee -a $S
