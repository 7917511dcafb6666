#!/bin/bash
# round 2, call AD: new prefetcher test + launch-policy knob sweep at the default batch (conv-engine ms per step)
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 600 python -m pytest tests/test_moving_digits.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -3
run() {
  tag=$1; shift
  env "$@" timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02ad_$tag.json 2> gpurun_out/r02ad_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02ad_$tag.json').read().strip().splitlines()[-1])
    k=d['roofline']['kernels']
    print('$tag', 'again', round(d['resident_again_ms_per_step'],2), 'fprop', round(k['igemm_fprop_kernel']['ms_per_step_in_kernel'],2), 'wgrad', round(k['igemm_wgrad_kernel']['ms_per_step_in_kernel'],2), 'halo', round(k['halo_fprop_kernel']['ms_per_step_in_kernel'],2), round(k['halo_wgrad_kernel']['ms_per_step_in_kernel'],2), 'conv', round(d['roofline']['conv_engine_all']['ms_per_step'],2))
except Exception as e:
    print('$tag failed', e)
PY
}
run default T2V_DUMMY=1
run persist256 T2V_FPROP_PERSIST_MIN_K=256
run persist1024 T2V_FPROP_PERSIST_MIN_K=1024
run wgbox16 T2V_WGRAD_MIN_BOXES=16
run wgbox64 T2V_WGRAD_MIN_BOXES=64
