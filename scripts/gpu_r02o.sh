#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_round2_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "batchnorm or fp32_storage" 2>&1 | grep -v "$F" | tail -4
timeout 600 python bench.py --no_cpu_baseline --no_library_baseline --steps 10 > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_o.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), round(d['e2e']['value']), d['roofline']['conv_engine_all'], d['gpu_launches'])
PY
