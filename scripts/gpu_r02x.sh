#!/bin/bash
mkdir -p gpurun_out
for b in 3072 4096; do
timeout -k 10 600 python bench.py --batch $b --steps 6 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02x_bench_b$b.json 2> gpurun_out/r02x_bench_b$b.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02x_bench_b$b.json').read().strip().splitlines()[-1])
    print($b, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['roofline']['conv_engine_all'], d['peak_mem_gb'], d['clocks'])
except Exception as e:
    print($b, 'failed', e)
PY
tail -2 gpurun_out/r02x_bench_b$b.err | cut -c1-300
done
