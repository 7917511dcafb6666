#!/bin/bash
# round 2, call AM: final evidence at HEAD: full GPU suite (-x), smoke, default bench (driver's flags, all arms)
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -25 > gpurun_out/r02am_pytest_gpu.log; tail -3 gpurun_out/r02am_pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/r02am_smoke.log
timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02am_bench_default.json 2> gpurun_out/r02am_bench_default.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02am_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fp32host', round(d['e2e_fp32_host']['value']), r['conv_engine_all']['frac'], r['traffic'], r['frac'], r['step_issued_frac_of_sustained_peak'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'], d['library_baseline']['best'])
PY
