#!/bin/bash
# round 2, call F: where does the fp32 mode's forward error come from?  + bench with the library comparator
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
T2V_FP32_SPLIT=3 timeout 300 python scripts/debug_fp32_conv.py 2>&1 | grep -v "$F" > gpurun_out/dbg_conv_split3.log
T2V_FP32_SPLIT=6 timeout 300 python scripts/debug_fp32_conv.py 2>&1 | grep -v "$F" > gpurun_out/dbg_conv_split6.log
cat gpurun_out/dbg_conv_split3.log gpurun_out/dbg_conv_split6.log
T2V_FP32_SPLIT=6 timeout 600 python scripts/debug_d_stages.py fp32 2>&1 | grep -v "$F" > gpurun_out/dbg_fp32_s6.log
grep -A 14 "=== generator" gpurun_out/dbg_fp32_s6.log; grep "stage\|=== level" gpurun_out/dbg_fp32_s6.log | head -24
timeout 900 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 1800 gpurun_out/bench_f.json; tail -5 gpurun_out/bench_f.err
