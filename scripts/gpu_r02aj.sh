#!/bin/bash
# round 2, call AJ: full GPU suite at HEAD (the driver's command) + smoke
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -25 > gpurun_out/r02aj_pytest_gpu.log; tail -3 gpurun_out/r02aj_pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -1 | tee gpurun_out/r02aj_smoke.log
