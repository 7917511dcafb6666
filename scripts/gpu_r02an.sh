#!/bin/bash
# round 2, call AN (2 GPUs): 2-rank parity test and a short 2-GPU bench line at HEAD
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 300 python -m pytest tests/test_nrank_parity.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -2
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 6 --warmup 5 --no_cpu_baseline --no_library_baseline > gpurun_out/r02an_bench_n2.json 2> gpurun_out/r02an_bench_n2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02an_bench_n2.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), 'again', round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'])
except Exception as e:
    print('failed', e)
PY
