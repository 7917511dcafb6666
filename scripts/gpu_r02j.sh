#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 900 python -m pytest tests/test_conv_engine_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | grep -v "$F" | tail -3
TOPN=400 timeout 600 python scripts/profile_shapes.py 2048 2>&1 | grep -v "$F" > gpurun_out/r02j_shapes_persist.txt
head -3 gpurun_out/r02j_shapes_persist.txt
timeout 600 python bench.py --no_cpu_baseline --no_library_baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['conv_engine_all'])
    for k,v in d['roofline']['kernels'].items(): print(k, v)
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_j.err').read()[-1500:])
PY
