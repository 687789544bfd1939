#!/usr/bin/env bash
# round 2, third GPU pass (1 GPU): fused large-batch CLIP, packed/bf16 targets, optim resume; bench lines for configs 1/2/4/5; CLIP sweep
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c3_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_optim_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -s \
   -k "fused or lse_merge or clip or optim or targets or golden" > gpurun_out/r2_c3_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|\[parity\] (targets|.*fused)|mbarrier" gpurun_out/r2_c3_tests.log | cut -c1-400 | tail -30 | tee -a $S
timeout 600 python scripts/clip_sweep.py > gpurun_out/r2_c3_clip_sweep.jsonl 2> gpurun_out/r2_c3_clip_sweep.err
echo "clip sweep rc=$?" | tee -a $S
cut -c1-700 gpurun_out/r2_c3_clip_sweep.jsonl | tee -a $S
tail -3 gpurun_out/r2_c3_clip_sweep.err | cut -c1-300 | tee -a $S
for c in gwilliams god brennan brennan49; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --max-seconds 800 > gpurun_out/r2_c3_bench_$c.json 2> gpurun_out/r2_c3_bench_$c.err
  echo "bench $c rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c3_bench_$c.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "targets", "graph_note", "loss", "gpu_launches", "step_frac_of_peak")},
          "roofline", r.get("frac"), r.get("avg_launch_ms"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["h2d_alone_ms"],
          "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("no bench line:", e)
PY
  tail -2 gpurun_out/r2_c3_bench_$c.err | cut -c1-300 | tee -a $S
done
