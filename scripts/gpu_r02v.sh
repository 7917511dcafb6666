#!/bin/bash
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout -k 10 900 python -m pytest tests/test_strided_tc_gpu.py tests/test_families.py tests/test_iteration_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v "$F" | tail -12
python scripts/families_once.py > gpurun_out/families_plain.log 2>&1 && tail -1 gpurun_out/families_plain.log &&
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02v_families_launches.csv python scripts/families_once.py > gpurun_out/ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/r02v_families_launches.csv 60 > gpurun_out/r02v_families_launches_summary.txt; head -24 gpurun_out/r02v_families_launches_summary.txt
