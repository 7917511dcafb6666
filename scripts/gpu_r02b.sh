#!/bin/bash
# round 2, GPU call B: failing tests with details, corrected bf16 floors, full per-shape conv profile, all-kernel launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_iteration_gpu.py tests/test_round2_gpu.py -q -s --tb=short \
  -k "fp32_mode or attention or conv_lstm" 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -200 > gpurun_out/pytest_b.log
for cfg in "cond 0" "cond 0.5" "uncond 0"; do
  set -- $cfg
  timeout 600 python scripts/bf16_floor.py $1 $2 cuda > gpurun_out/r02_bf16_floor_$1_g$2.json 2> gpurun_out/floor_$1_$2.err
done
timeout 600 python scripts/bf16_floor_families.py cuda > gpurun_out/r02_bf16_floor_families.json 2> gpurun_out/floor_families.err
TOPN=400 timeout 900 python scripts/profile_shapes.py 2048 2>&1 | grep -v "Warn\|Consider\|run_backward" > gpurun_out/r02b_conv_shapes_b2048_all.txt
python scripts/iter_once.py --batch 1024 > gpurun_out/iter_plain.log 2>&1 &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02b_launches_b1024.csv python scripts/iter_once.py --batch 1024 > gpurun_out/ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/r02b_launches_b1024.csv 80 > gpurun_out/r02b_launches_b1024_summary.txt
grep -n "passed\|failed" gpurun_out/pytest_b.log | tail -3
cat gpurun_out/r02_bf16_floor_*.json
head -30 gpurun_out/r02b_launches_b1024_summary.txt
