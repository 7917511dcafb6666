#!/bin/bash
# round 2, call AI: uint8 frames fused into the graph's input fill: tests, then the default bench line
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_graphed_step_gpu.py tests/test_kernels_gpu.py tests/test_moving_digits.py -q -m gpu -p no:cacheprovider --tb=short -k "graph or prefetcher or uint8 or cli or moving" 2>&1 | grep -v "$F" | tail -4
timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no_library_baseline > gpurun_out/r02ai_bench_default.json 2> gpurun_out/r02ai_bench_default.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02ai_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'fp32host', round(d['e2e_fp32_host']['value']), r['conv_engine_all']['frac'], r['frac'], d['clocks'])
PY
tail -2 gpurun_out/r02ai_bench_default.err | cut -c1-200
