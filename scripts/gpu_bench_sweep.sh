#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/bench_sweep.log
for b in ${BATCHES:-8 64 256 512 1024}; do
  echo "=== batch $b" | tee -a gpurun_out/bench_sweep.log
  timeout -k 10 900 python bench.py --batch $b --steps ${STEPS:-6} --warmup 3 --no_cpu_baseline 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -3 | tee -a gpurun_out/bench_sweep.log
done
