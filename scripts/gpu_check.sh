#!/usr/bin/env bash
# Runs the GPU test tiers in separate processes (a faulted CUDA context must not hide later results) and a short bench.
# Usage (on the GPU box, from the repo root):  bash scripts/gpu_check.sh [stage ...]
set -u
mkdir -p gpurun_out
run() {  # name, timeout, command...
  local name="$1" to="$2"; shift 2
  echo "=== ${name}" | tee -a gpurun_out/summary.txt
  timeout "${to}" "$@" > "gpurun_out/${name}.log" 2>&1
  local rc=$?
  echo "rc=${rc} $(tail -n 1 gpurun_out/${name}.log | cut -c1-300)" | tee -a gpurun_out/summary.txt
  grep -E "^(FAILED|ERROR)|\[parity\]|megb200: mbarrier" "gpurun_out/${name}.log" | cut -c1-400 | head -40 | tee -a gpurun_out/summary.txt
  return ${rc}
}
stages="${*:-ops_simt tc_tapgemm tc_wgrad tc_clip parity_fp32 parity_bf16_simt parity_bf16 parity_misc bench}"
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv | tee -a gpurun_out/summary.txt
PT="python -m pytest -q -m gpu -p no:cacheprovider"
for s in ${stages}; do
  case "${s}" in
    ops_simt)         run ops_simt 900 ${PT} tests/test_ops_gpu.py -k "not tc" ;;
    tc_tapgemm)       run tc_tapgemm 300 ${PT} tests/test_ops_gpu.py -k "tapgemm and tc" ;;
    tc_all)           run tc_all 600 ${PT} tests/test_ops_gpu.py -k "tc" ;;
    tc_halo1)         MEGB200_TAPGEMM_HALO=1 run tc_halo1 600 ${PT} tests/test_ops_gpu.py -k "tapgemm and tc" ;;
    tc_pair)          MEGB200_TAPGEMM_2CTA_MIN_ITEMS=1 MEGB200_TAPGEMM_PAIR_HALO=0 run tc_pair 600 ${PT} tests/test_ops_gpu.py -k "tapgemm and tc" ;;
    tc_pairhalo)      MEGB200_TAPGEMM_2CTA_MIN_ITEMS=1 run tc_pairhalo 600 ${PT} tests/test_ops_gpu.py -k "tapgemm and tc" ;;
    tc_1cta)          MEGB200_TAPGEMM_1CTA=1 run tc_1cta 600 ${PT} tests/test_ops_gpu.py -k "tapgemm and tc" ;;
    parity_pair)      MEGB200_TAPGEMM_2CTA_MIN_ITEMS=1 run parity_pair 600 ${PT} -s tests/test_parity_gpu.py -k "bf16 and not bf16_simt and not full_size and not default" ;;
    tc_wgrad)         run tc_wgrad 300 ${PT} tests/test_ops_gpu.py -k "wgrad and tc" ;;
    tc_clip)          run tc_clip 300 ${PT} tests/test_ops_gpu.py -k "clip and tc" ;;
    parity_fp32)      run parity_fp32 900 ${PT} -s tests/test_parity_gpu.py -k "fp32 and not full_size" ;;
    parity_bf16_simt) run parity_bf16_simt 900 ${PT} -s tests/test_parity_gpu.py -k "bf16_simt" ;;
    parity_bf16)      run parity_bf16 600 ${PT} -s tests/test_parity_gpu.py -k "bf16 and not bf16_simt and not full_size and not default" ;;
    parity_misc)      run parity_misc 900 ${PT} -s tests/test_parity_gpu.py -k "full_size or default or native" ;;
    bench)            run bench 900 python bench.py --steps 5 --warmup 3 ;;
    bench_b32)        run bench_b32 600 python bench.py --steps 3 --warmup 3 --batch 32 --no-cpu-baseline ;;
    smoke)            run smoke 300 python __graft_entry__.py smoke ;;
    ncu_launches)     run ncu_plain 300 python bench.py --ncu-step --warmup 3 &&
                      run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
                          --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --warmup 3 ;;
    ncu_full)         run ncu_plain2 300 python bench.py --ncu-step --warmup 3 &&
                      run ncu_full 1200 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
                          -k regex:"${NCU_KERNEL:-tc_tapgemm}" -s "${NCU_SKIP:-8}" -c "${NCU_COUNT:-2}" -f -o gpurun_out/prof_${NCU_TAG:-tapgemm} \
                          python bench.py --ncu-step --warmup 3 ;;
    all_gpu_tests)    run all_gpu_tests 1800 ${PT} tests ;;
    *) echo "unknown stage ${s}" ;;
  esac
done
echo "=== done"; cat gpurun_out/summary.txt | tail -n 5 > /dev/null
