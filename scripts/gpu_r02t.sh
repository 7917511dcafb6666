#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_moving_digits.py tests/test_round2_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --tb=short -s -k "moving or u8 or caption_pretraining or prefetcher" 2>&1 | grep -v "$F" | grep "passed\|failed\|FAILED\|Error\|deviations\|assert" | cut -c1-600
