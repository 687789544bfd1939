#!/usr/bin/env bash
# round 2, 2-GPU pass: NCCL + peer-gather DP parity test, 2-GPU bench (peer gather / nccl gather / no early bucket), and
# the rest of the 1-GPU suite that the first pass did not reach
set -u
mkdir -p gpurun_out
S=gpurun_out/r2_c2_summary.txt
: > $S
nvidia-smi topo -m > gpurun_out/r2_c2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -p no:cacheprovider -s > gpurun_out/r2_c2_dp.log 2>&1
echo "dp tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|\[parity\]|Error" gpurun_out/r2_c2_dp.log | cut -c1-500 | tail -20 | tee -a $S
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_c2_parity.log 2>&1
echo "parity tests rc=$?" | tee -a $S
grep -E "^(FAILED|ERROR)|passed|failed|\[parity\] gwill" gpurun_out/r2_c2_parity.log | cut -c1-700 | tail -20 | tee -a $S
run_bench() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_c2_bench_${tag}.json 2> gpurun_out/r2_c2_bench_${tag}.err
  echo "bench ${tag} rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("gpurun_out/r2_c2_bench_${tag}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_eager", "dp_target_gather", "dp_target_gather_note", "graph_note", "loss", "gpu_launches")}, d["roofline"]["avg_launch_ms"], d["e2e"]["value"])
except Exception as e:
    print("no bench line:", e)
PY
}
run_bench peer MEGB200_DP_GATHER=peer
run_bench nccl MEGB200_DP_GATHER=nccl
run_bench peer_noearly MEGB200_DP_GATHER=peer MEGB200_DP_EARLY_BUCKET=0
tail -5 gpurun_out/r2_c2_bench_peer.err | cut -c1-300 | tee -a $S
