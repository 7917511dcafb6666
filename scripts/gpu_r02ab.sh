#!/bin/bash
# round 2, call AB (8 GPUs): the driver's own command lines at the default batch (b = 4096 / GPU): N = 8, then N = 1
mkdir -p gpurun_out
free -g | head -2
timeout -k 10 870 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02ab_bench_n8.json 2> gpurun_out/r02ab_bench_n8.err
echo "n8 rc=$?"
timeout -k 10 870 python bench.py --gpus 1 --steps 20 --warmup 5 --no_library_baseline > gpurun_out/r02ab_bench_n1.json 2> gpurun_out/r02ab_bench_n1.err
echo "n1 rc=$?"
python - <<PY
import json
for n in (8, 1):
    try:
        d=json.loads(open('gpurun_out/r02ab_bench_n%d.json' % n).read().strip().splitlines()[-1])
        print(n, round(d['value']), round(d['ms_per_step'],2), 'again', round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fp32host', round(d['e2e_fp32_host']['value']), d['roofline']['traffic'], d['roofline']['frac'], d['clocks'], d.get('cpu_baseline'))
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/r02ab_bench_n8.err | cut -c1-300
