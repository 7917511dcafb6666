#!/bin/bash
# One GPU call: iteration parity, per-shape conv profile, default bench, H2D probe.
B=${B:-1024}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_iteration_gpu.py tests/test_graphed_step_gpu.py -q -x 2>&1 | tail -3
timeout -k 10 600 python scripts/profile_shapes.py $B 2>&1 | grep -v "Warn\|Consider\|run_backward" | head -70
timeout -k 10 900 python bench.py --batch $B --steps 6 --warmup 3 --no_cpu_baseline 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -1 | tee gpurun_out/bench_b$B.log
timeout -k 10 300 python scripts/h2d_probe.py 2>&1 | tail -5 | tee gpurun_out/h2d_probe.log
