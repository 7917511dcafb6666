#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
for i in 1 2 3; do
timeout 900 python -m pytest tests/test_families.py tests/test_round2_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider -k "tgan or tcwyt or fused_lstm" 2>&1 | grep -v "$F" | grep "tgan gpu\|tcwyt gpu\|passed\|failed\|^FAILED\|Error" | cut -c1-420
done
