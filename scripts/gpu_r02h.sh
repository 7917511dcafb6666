#!/bin/bash
# round 2, call H: persistent igemm fprop -- correctness (conv engine tests) and A/B per-shape timing at b = 2048
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 900 python -m pytest tests/test_conv_engine_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | grep -v "$F" | tail -15 > gpurun_out/pytest_h.log
tail -5 gpurun_out/pytest_h.log
T2V_FPROP_PERSIST=0 TOPN=400 timeout 600 python scripts/profile_shapes.py 2048 2>&1 | grep -v "$F" > gpurun_out/r02h_shapes_nopersist.txt
TOPN=400 timeout 600 python scripts/profile_shapes.py 2048 2>&1 | grep -v "$F" > gpurun_out/r02h_shapes_persist.txt
head -3 gpurun_out/r02h_shapes_nopersist.txt; head -3 gpurun_out/r02h_shapes_persist.txt
timeout 600 python bench.py --no_cpu_baseline --no_library_baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['conv_engine_all'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_h.err').read()[-1500:])
PY
