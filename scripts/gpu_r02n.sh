#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_iteration_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider -k "caption or full_iteration_on_b200 or bf16_attention" 2>&1 | grep -v "$F" | grep "passed\|failed\|^FAILED\|Error\|deviations" | cut -c1-500
python scripts/iter_once.py --batch 2048 > gpurun_out/iter_plain.log 2>&1 &&
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02n_launches_b2048.csv python scripts/iter_once.py --batch 2048 > gpurun_out/ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/r02n_launches_b2048.csv 90 > gpurun_out/r02n_launches_b2048_summary.txt
head -75 gpurun_out/r02n_launches_b2048_summary.txt
