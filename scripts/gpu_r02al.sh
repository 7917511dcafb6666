#!/bin/bash
# round 2, call AL: fused generator attention on maps beyond one CTA (t2v_attention_bwd_large): kernel tests, the
# 128x128x32 iteration, then the configs[4] bench line
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_round2_gpu.py tests/test_iteration_gpu.py -q -m gpu -p no:cacheprovider --tb=short -k "attention or config5" 2>&1 | grep -v "$F" | tail -8
timeout -k 10 900 python bench.py --res 128 --batch 128 --steps 8 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02al_bench_res128_b128.json 2> gpurun_out/r02al_bench_res128_b128.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02al_bench_res128_b128.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), r['conv_engine_all'], d['peak_mem_gb'])
PY
tail -2 gpurun_out/r02al_bench_res128_b128.err | cut -c1-200
