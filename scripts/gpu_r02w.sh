#!/bin/bash
# round 2, call W (8 GPUs): fixed window-route tests, then the 8-GPU and 1-GPU bench lines on the same box
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 600 python -m pytest tests/test_strided_tc_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --tb=short -k "engine or gconv" 2>&1 | grep -v "$F" | tail -4
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err
python - <<PY
import json
for n in (8, 1):
    try:
        d=json.loads(open('gpurun_out/r02w_bench_n%d.json' % n).read().strip().splitlines()[-1])
        print(n, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fp32host', round(d['e2e_fp32_host']['value']), d['clocks'])
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/r02w_bench_n8.err | cut -c1-300
