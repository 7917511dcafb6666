#!/bin/bash
# round 2, call U: full GPU suite at HEAD, smoke, families launch list (ncu), default bench
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -25 > gpurun_out/r02u_pytest_gpu.log; tail -4 gpurun_out/r02u_pytest_gpu.log
timeout -k 10 600 python -m pytest tests/test_families.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep "gpu:\|gpu fp32" | cut -c1-400 | tee gpurun_out/r02u_families.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/r02u_smoke.log
python scripts/families_once.py > gpurun_out/families_plain.log 2>&1 && tail -1 gpurun_out/families_plain.log &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02u_families_launches.csv python scripts/families_once.py > gpurun_out/ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/r02u_families_launches.csv 60 > gpurun_out/r02u_families_launches_summary.txt; head -30 gpurun_out/r02u_families_launches_summary.txt
timeout -k 10 900 python bench.py --no_library_baseline > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fp32host', round(d['e2e_fp32_host']['value']), r['conv_engine_all'], r['traffic'], r['step_issued_frac_of_sustained_peak'], d['gpu_launches'])
PY
