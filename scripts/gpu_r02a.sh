#!/bin/bash
# round 2, GPU call A: bf16 floors of the oracle (autocast vs fp32), the whole GPU test suite, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
for cfg in "cond 0" "cond 0.5" "uncond 0"; do
  set -- $cfg
  timeout 600 python scripts/bf16_floor.py $1 $2 cuda > gpurun_out/r02_bf16_floor_$1_g$2.json 2> gpurun_out/floor_$1_$2.err
done
timeout 600 python scripts/bf16_floor_families.py cuda > gpurun_out/r02_bf16_floor_families.json 2> gpurun_out/floor_families.err
timeout 2400 python -m pytest tests -m gpu -q -rf 2>&1 | tail -200 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/r02_bf16_floor_*.json
tail -c 1500 gpurun_out/bench_a.json
