#!/bin/bash
# One GPU call: kernel-level parity, conv engine, full-iteration parity.  Each group in its own process.
mkdir -p gpurun_out
for t in tests/test_kernels_gpu.py tests/test_conv_engine_gpu.py tests/test_iteration_gpu.py; do
  echo "=== $t" | tee -a gpurun_out/groups.log
  timeout -k 10 900 python -m pytest $t -q -m gpu -s 2>&1 | grep -v "Warn\|Consider\|out\[" | tail -40 | tee -a gpurun_out/groups.log
done
