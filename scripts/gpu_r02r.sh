#!/bin/bash
# round 2, call R: strided convolutions on the engine + families, and the fp32 iteration repeated (flake check)
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_strided_tc_gpu.py tests/test_families.py -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | grep -v "$F" | tail -30
for i in 1 2 3; do
timeout -k 10 600 python -m pytest tests/test_iteration_gpu.py -q -m gpu -p no:cacheprovider --tb=line -s -k "fp32_mode" 2>&1 | grep "deviations\|passed\|failed" | cut -c1-700
done
