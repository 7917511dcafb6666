#!/bin/bash
# round 2, call S (2 GPUs): 2-rank parity test on the GPU, new caption pre-training test, 2-GPU bench line
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_nrank_parity.py tests/test_round2_gpu.py -q -m gpu -p no:cacheprovider --tb=short -s -k "two_rank or caption_pretraining" 2>&1 | grep -v "$F" | grep "passed\|failed\|FAILED\|Error\|deviations\|assert" | cut -c1-600
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no_cpu_baseline --no_library_baseline > gpurun_out/r02s_bench_n2.json 2> gpurun_out/r02s_bench_n2.err
tail -c 1200 gpurun_out/r02s_bench_n2.json; tail -3 gpurun_out/r02s_bench_n2.err
