#!/bin/bash
# round 2, call AL: fused generator attention on maps beyond one CTA (t2v_attention_bwd_large): kernel tests, the
# 128x128x32 iteration, then the configs[4] bench line
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
true
timeout -k 10 900 python bench.py --res 128 --batch 512 --steps 8 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02al_bench_res128_b512.json 2> gpurun_out/r02al_bench_res128_b512.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02al_bench_res128_b512.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), r['conv_engine_all'], d['peak_mem_gb'])
PY
tail -2 gpurun_out/r02al_bench_res128_b512.err | cut -c1-200
