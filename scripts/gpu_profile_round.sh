#!/bin/bash
# One GPU call: iteration parity, per-shape conv profile, ncu launch list (all kernels), bench.
B=${B:-1024}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_iteration_gpu.py tests/test_graphed_step_gpu.py -q -x 2>&1 | tail -3
timeout -k 10 600 python scripts/profile_shapes.py $B 2>&1 | grep -v "Warn\|Consider\|run_backward" | head -50
timeout -k 10 900 python bench.py --batch $B --steps 6 --warmup 3 --no_cpu_baseline 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -1 | tee gpurun_out/bench_b$B.log
python scripts/iter_once.py --batch ${NB:-256} > gpurun_out/iter_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_b${NB:-256}.csv python scripts/iter_once.py --batch ${NB:-256} > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_b${NB:-256}.csv
