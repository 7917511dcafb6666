#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 600 python scripts/debug_d_stages.py fp32 2>&1 | grep -v "$F" > gpurun_out/dbg_fp32.log
timeout 600 python scripts/debug_d_stages.py fp32 perturb 2>&1 | grep -v "$F" > gpurun_out/dbg_fp32_perturb.log
timeout 600 python scripts/debug_d_stages.py bf16 perturb 2>&1 | grep -v "$F" > gpurun_out/dbg_bf16_perturb.log
tail -60 gpurun_out/dbg_fp32.log
