#!/bin/bash
# round 2, call AA: DRAM traffic + time of every kernel of one iteration at the default bench batch (b = 4096):
# source of roofline.traffic (profiles/r02_traffic_b4096.json) and of the launch list at that batch
mkdir -p gpurun_out
python scripts/iter_once.py --batch 4096 > gpurun_out/iter_plain.log 2>&1 || { tail -5 gpurun_out/iter_plain.log; exit 1; }
timeout -k 10 1800 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_traffic_b4096.csv python scripts/iter_once.py --batch 4096 > gpurun_out/ncu.log 2>&1
python scripts/agg_traffic.py gpurun_out/r02_traffic_b4096.csv > gpurun_out/r02_traffic_b4096.json
python scripts/agg_launches.py gpurun_out/r02_traffic_b4096.csv 90 > gpurun_out/r02aa_launches_b4096_summary.txt
head -40 gpurun_out/r02aa_launches_b4096_summary.txt
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02aa_bench_default.json 2> gpurun_out/r02aa_bench_default.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02aa_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],2), round(d['resident_again_ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fp32host', round(d['e2e_fp32_host']['value']), r['conv_engine_all'], r['traffic'], r['frac'], d['gpu_launches'], d['config'])
PY
