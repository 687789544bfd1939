#!/usr/bin/env bash
# A/B bench of the tap-GEMM kernel variants (env toggles of csrc/tc_gemm.cu); prints one line per variant.
run() {
  local tag="$1"; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('VARIANT $tag ms/step=%.3f e2e=%.3f host=%.2f tapgemm=%.1f TF/s (%.3f ms/launch) wgrad=%.1f TF/s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['host_enqueue_ms_per_step'], r['achieved'], r['avg_launch_ms'], r['wgrad']['achieved']))"
}
run default       A=1
run pair_nohalo   MEGB200_TAPGEMM_PAIR_HALO=0
run pair_ragged   MEGB200_TAPGEMM_RAGGED_N=1
run 1cta_tma      MEGB200_TAPGEMM_1CTA=1
run 1cta_direct   MEGB200_TAPGEMM_1CTA=1 MEGB200_TAPGEMM_DIRECT_EPI=1
run 1cta_halo     MEGB200_TAPGEMM_1CTA=1 MEGB200_TAPGEMM_HALO=1
