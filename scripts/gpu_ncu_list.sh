#!/bin/bash
mkdir -p gpurun_out
python scripts/iter_once.py --batch 64 > gpurun_out/iter_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_b64.csv python scripts/iter_once.py --batch 64 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/iter_plain.log; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches_b64.csv
