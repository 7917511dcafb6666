#!/bin/bash
# round 2, call AC: full GPU suite at HEAD (the driver's command, with -x) and smoke
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -25 > gpurun_out/r02ac_pytest_gpu.log; tail -6 gpurun_out/r02ac_pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/r02ac_smoke.log
