#!/usr/bin/env bash
# round 2: final-code bench lines of the other BASELINE configs (1 GPU; the library travels prebuilt)
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c17_summary.txt
: > $S
run() {
  c=$1; shift
  timeout 75 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --max-seconds 70 "$@" > gpurun_out/r2_c17_bench_$c.json 2> gpurun_out/r2_c17_bench_$c.err
  echo "bench $c rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c17_bench_$c.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "loss", "gpu_launches")}, "frac", r.get("frac"),
          "e2e", d["e2e"]["value"], "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("no bench line:", e)
PY
  tail -1 gpurun_out/r2_c17_bench_$c.err | cut -c1-200 | tee -a $S
}
run brennan49
run brennan
run god --no-cpu-baseline
