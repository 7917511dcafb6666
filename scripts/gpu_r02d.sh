#!/bin/bash
# round 2, call D: full GPU suite after the fp32 im2col fix and floor-based bars; smoke; default bench
mkdir -p gpurun_out
F='Warn\|Consider\|run_backward'
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -v "$F" > gpurun_out/pytest_d.log
grep -n "passed\|failed" gpurun_out/pytest_d.log | tail -3
grep -n "^FAILED\|deviations\|tgan gpu\|tcwyt gpu\|autocast floor\|product:\|generator attention" gpurun_out/pytest_d.log | cut -c1-600
timeout 600 python __graft_entry__.py smoke 2>&1 | grep -v "$F" | tail -3 > gpurun_out/smoke_d.log; cat gpurun_out/smoke_d.log
timeout 900 python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 3000 gpurun_out/bench_d.json
