#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
for st in 6 10; do
timeout -k 10 600 python bench.py --batch 4096 --steps $st --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02z_b4096_s$st.json 2> gpurun_out/r02z_b4096_s$st.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02z_b4096_s$st.json').read().strip().splitlines()[-1])
    print('steps=$st', round(d['value']), round(d['ms_per_step'],2), 'again', round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'])
except Exception as e:
    print('failed', e)
PY
done
