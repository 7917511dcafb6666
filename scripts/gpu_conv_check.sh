#!/bin/bash
# Runs the conv-engine GPU tests group by group (separate processes, so a trap in one group does not
# poison the others) and collects everything under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/conv_engine.log
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "test_fprop and simt and not epilogue" "test_fprop and tc and not epilogue" "test_fprop_epilogue" "test_dgrad" "test_wgrad and simt" "test_wgrad and tc" "test_simt_odd" "test_perf_probe"; do
  echo "=== $grp" | tee -a gpurun_out/conv_groups.log
  timeout -k 10 300 python -m pytest tests/test_conv_engine_gpu.py -q -x -k "$grp" 2>&1 | tail -15 | tee -a gpurun_out/conv_groups.log
done
cat gpurun_out/conv_engine.log
