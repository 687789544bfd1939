#!/usr/bin/env bash
# round 2: final default bench line incl. the optional resident-targets leg (1 GPU)
set -u
mkdir -p gpurun_out
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || echo "BUILD FAILED"
timeout 600 python bench.py --e2e-resident > gpurun_out/r2_c16_bench.json 2> gpurun_out/r2_c16_bench.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_c16_bench.json").read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "ms_per_step_unchanged_loop_module_graphs", "loss")})
print("roofline", {k: r.get(k) for k in ("achieved", "frac", "traffic", "ncu_tensor_pipe_active_pct", "traffic_source")})
print("e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "h2d_bytes_per_step", "resident_targets")})
print("parity", (d.get("parity") or {}).get("z_rel"))
PY
tail -2 gpurun_out/r2_c16_bench.err | cut -c1-300
