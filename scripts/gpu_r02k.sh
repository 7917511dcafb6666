#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 0 1; do
T2V_FPROP_PERSIST=$p timeout 600 python bench.py --no_cpu_baseline --no_library_baseline --steps 10 > gpurun_out/bench_k$p.json 2> gpurun_out/bench_k.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_k$p.json').read().strip().splitlines()[-1])
print("persist=$p", round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), round(d['e2e']['value']), d['roofline']['conv_engine_all'], d['clocks'])
PY
done
