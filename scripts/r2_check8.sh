#!/usr/bin/env bash
# round 2, N-GPU pass: weak-scaling bench at N GPUs (peer copy-engine gather) with a per-kernel timeline of one step on rank 0
set -u
N=${1:-8}
mkdir -p gpurun_out
S=gpurun_out/r2_c8_summary_n$N.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
nvidia-smi topo -m > gpurun_out/r2_c8_topo.txt 2>&1
timeout ${MAXT:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --max-seconds ${MAXS:-300} --no-e2e --timeline gpurun_out/r2_timeline_n$N.json \
    > gpurun_out/r2_c8_bench_n$N.json 2> gpurun_out/r2_c8_bench_n$N.err
echo "bench n$N rc=$?" | tee -a $S
python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c8_bench_n$N.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "ms_per_step_eager", "dp_target_gather", "dp_target_gather_note", "graph_note", "loss", "gpu_launches")},
          "avg_launch_ms", r.get("avg_launch_ms"), "frac", r.get("frac"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    t = json.load(open("gpurun_out/r2_timeline_n$N.json"))
    print("timeline span_us", t["span_us"])
    for k in t["by_kernel"][:14]:
        print("  ", k)
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/r2_c8_bench_n$N.err | cut -c1-300 | tee -a $S
