#!/bin/bash
# round 2, call Q: full GPU suite at HEAD (no -x: every failure is listed), smoke, default bench with both baselines
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -40 > gpurun_out/r02q_pytest_gpu.log; tail -5 gpurun_out/r02q_pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -3 | tee gpurun_out/r02q_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r02q_bench_default.json 2> gpurun_out/r02q_bench_default.err; tail -c 600 gpurun_out/r02q_bench_default.json
