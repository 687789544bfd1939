#!/usr/bin/env bash
# round 2, fourth GPU pass (1 GPU): full GPU suite on the rebuilt library, CLIP sweep, god bench parity
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c4_summary.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; tail -5 gpurun_out/r2_build.log | tee -a $S; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_c4_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|mbarrier|\[parity\] (targets|.*fused|.*pooled|.*dilated|classifier)" gpurun_out/r2_c4_tests.log | cut -c1-300 | tail -60 | tee -a $S
timeout 600 python scripts/clip_sweep.py > gpurun_out/r2_c4_clip_sweep.jsonl 2> gpurun_out/r2_c4_clip_sweep.err
echo "clip sweep rc=$?" | tee -a $S
cut -c1-900 gpurun_out/r2_c4_clip_sweep.jsonl | tee -a $S
tail -3 gpurun_out/r2_c4_clip_sweep.err | cut -c1-300 | tee -a $S
for c in god; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --max-seconds 800 > gpurun_out/r2_c4_bench_$c.json 2> gpurun_out/r2_c4_bench_$c.err
  echo "bench $c rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c4_bench_$c.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "loss")}, "parity", d.get("parity"))
except Exception as e:
    print("no bench line:", e)
PY
done
