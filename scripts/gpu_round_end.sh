#!/bin/bash
# Round-end evidence in one GPU call: full GPU test-suite, smoke(), default bench (+ reference arm), per-shape conv
# profile.  (ncu captures: scripts/gpu_ncu_full.sh, scripts/gpu_ncu_list.sh)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -2 | tee gpurun_out/smoke.log
timeout -k 10 900 python bench.py 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -1 | tee gpurun_out/bench_default.json | cut -c1-200
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
timeout -k 10 600 python scripts/profile_shapes.py 2048 2>&1 | grep -v "Warn\|Consider\|run_backward" > gpurun_out/shapes_b2048.txt; head -3 gpurun_out/shapes_b2048.txt
