#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
T2V_DEBUG_LAUNCH=1 T2V_FPROP_PERSIST=1 timeout 300 python scripts/perf_shapes.py P1 2>&1 | grep -v "$F" | grep "igemm_fprop\|fprop" | awk '!seen[$0]++' | head -42 > gpurun_out/perf_p1.log
T2V_FPROP_PERSIST=1 T2V_FPROP_2CTA=0 timeout 300 python scripts/perf_shapes.py P1big 2>&1 | grep -v "$F" | grep "^P1big" | head -21 > gpurun_out/perf_p1big.log
cat gpurun_out/perf_p1.log gpurun_out/perf_p1big.log | cut -c1-200
