#!/bin/bash
mkdir -p gpurun_out
for nc in 1 0; do
T2V_BENCH_NO_CLOCKS=$nc timeout -k 10 600 python bench.py --batch 4096 --steps 8 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02y_b4096_nc$nc.json 2> gpurun_out/r02y_b4096_nc$nc.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02y_b4096_nc$nc.json').read().strip().splitlines()[-1])
    print('noclocks=$nc', round(d['value']), round(d['ms_per_step'],2), 'again', round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'])
except Exception as e:
    print('failed', e)
PY
done
