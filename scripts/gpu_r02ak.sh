#!/bin/bash
# round 2, call AK: BASELINE configs[4] (TGANv2-cond 128x128x32) on one GPU
mkdir -p gpurun_out
timeout -k 10 900 python bench.py --res 128 --batch 128 --steps 8 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02ak_bench_res128_b128.json 2> gpurun_out/r02ak_bench_res128_b128.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02ak_bench_res128_b128.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), r['conv_engine_all'], r['frac'], r['step_nominal_frac_of_sustained_peak'], r['step_issued_frac_of_sustained_peak'], d['peak_mem_gb'], d['config']['workload'])
PY
tail -2 gpurun_out/r02ak_bench_res128_b128.err | cut -c1-300
