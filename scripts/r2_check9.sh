#!/usr/bin/env bash
# round 2, N-GPU pass 2: bench lines of other configs / tunings at N GPUs.  usage: r2_check9.sh N "tag:config:ENV=..,ENV=.." ...
set -u
N=$1; shift
mkdir -p gpurun_out
S=gpurun_out/r2_c9_summary_n$N.txt
: > $S
bash meg-decoding_b200/csrc/build.sh > gpurun_out/r2_build.log 2>&1 || { echo "BUILD FAILED" | tee -a $S; }
for spec in "$@"; do
  tag=${spec%%:*}; rest=${spec#*:}; cfg=${rest%%:*}; envs=${rest#*:}; [ "$envs" = "$rest" ] && envs=""
  env $(echo $envs | tr ',' ' ') timeout ${MAXT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $N --config $cfg --steps 10 --warmup 3 --max-seconds ${MAXS:-240} --no-e2e --timeline gpurun_out/r2_timeline_${tag}_n$N.json \
      > gpurun_out/r2_c9_bench_${tag}_n$N.json 2> gpurun_out/r2_c9_bench_${tag}_n$N.err
  echo "bench $tag ($cfg, $envs) n$N rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c9_bench_${tag}_n$N.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "ms_per_step_eager", "dp_target_gather", "graph_note", "loss", "gpu_launches")},
          "avg_launch_ms", r.get("avg_launch_ms"), "e2e", d["e2e"]["value"])
    t = json.load(open("gpurun_out/r2_timeline_${tag}_n$N.json"))
    print("timeline span_us", t["span_us"], [k for k in t["by_kernel"] if "nccl" in k[0] or "PtoP" in k[0] or "clip" in k[0]][:8])
except Exception as e:
    print("no bench line:", e)
PY
  tail -2 gpurun_out/r2_c9_bench_${tag}_n$N.err | cut -c1-300 | tee -a $S
done
