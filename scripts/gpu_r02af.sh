#!/bin/bash
# round 2, call AF: weight-gradient CTA order (tap fastest vs split fastest): engine tests, then A/B at the default batch
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 600 python -m pytest tests/test_conv_engine_gpu.py tests/test_strided_tc_gpu.py tests/test_graphed_step_gpu.py -q -m gpu -p no:cacheprovider --tb=short -k "wgrad" 2>&1 | grep -v "$F" | tail -3
run() {
  tag=$1; shift
  env "$@" timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02af_$tag.json 2> gpurun_out/r02af_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02af_$tag.json').read().strip().splitlines()[-1])
    k=d['roofline']['kernels']
    print('$tag', 'again', round(d['resident_again_ms_per_step'],2), 'fprop', round(k['igemm_fprop_kernel']['ms_per_step_in_kernel'],2), 'wgrad', round(k['igemm_wgrad_kernel']['ms_per_step_in_kernel'],2), round(k['igemm_wgrad_kernel']['achieved'],1), 'halo', round(k['halo_fprop_kernel']['ms_per_step_in_kernel'],2), round(k['halo_wgrad_kernel']['ms_per_step_in_kernel'],2), 'conv', round(d['roofline']['conv_engine_all']['ms_per_step'],2), round(d['roofline']['conv_engine_all']['frac'],4))
except Exception as e:
    print('$tag failed', e)
PY
}
run tapfast1 T2V_WGRAD_TAP_FAST=1
run tapfast0 T2V_WGRAD_TAP_FAST=0
run tapfast1b T2V_WGRAD_TAP_FAST=1
