#!/usr/bin/env bash
# round 2, fifth GPU pass (1 GPU): full GPU suite (CTA-pair weight gradient, ModuleGraphs), config-2 bench with the
# single-CTA weight-gradient kernel for comparison
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c5_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_c5_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier|\[parity\] (module|gwilliams B=256)" gpurun_out/r2_c5_tests.log | cut -c1-400 | tail -40 | tee -a $S
bench() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 900 python bench.py --steps 10 --warmup 3 --max-seconds 800 > gpurun_out/r2_c5_bench_$tag.json 2> gpurun_out/r2_c5_bench_$tag.err
  echo "bench $tag rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c5_bench_$tag.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "module_graphs_note", "graph_note", "loss", "gpu_launches", "step_frac_of_peak", "host_enqueue_ms_per_step")},
          "roofline", r.get("frac"), r.get("avg_launch_ms"), "wgrad", r.get("wgrad"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", (d.get("parity") or {}).get("z_rel"))
except Exception as e:
    print("no bench line:", e)
PY
  tail -2 gpurun_out/r2_c5_bench_$tag.err | cut -c1-300 | tee -a $S
}
bench pair MEGB200_WGRAD_PAIR=1
bench single MEGB200_WGRAD_PAIR=0
