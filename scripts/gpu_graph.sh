#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_graphed_step_gpu.py -q -s -x 2>&1 | grep -v "Warn\|Consider\|run_backward" | tail -30 | tee gpurun_out/graph_test.log
