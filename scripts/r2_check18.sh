#!/usr/bin/env bash
# round 2: 2-GPU data-parallel tests after the PeerGather constructor change (library travels prebuilt)
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_c18_dp.log 2>&1
echo "dp tests rc=$?"
grep -E "passed|failed|FAILED|ERROR|gather" gpurun_out/r2_c18_dp.log | cut -c1-300 | tail -12
