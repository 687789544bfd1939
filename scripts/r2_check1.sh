#!/usr/bin/env bash
# round 2, first GPU pass: the whole GPU suite (new product-kernel parity tests), smoke, a short bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/r2_c1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider -s > gpurun_out/r2_c1_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/r2_c1_summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|\[parity\]" gpurun_out/r2_c1_tests.log | cut -c1-600 | tail -60 | tee -a gpurun_out/r2_c1_summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_c1_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r2_c1_summary.txt
tail -3 gpurun_out/r2_c1_smoke.log | cut -c1-400 | tee -a gpurun_out/r2_c1_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2_c1_summary.txt
tail -c 3000 gpurun_out/r2_c1_bench.json | tee -a gpurun_out/r2_c1_summary.txt
