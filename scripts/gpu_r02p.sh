#!/bin/bash
# round 2, call P: ncu --set full of the conv / attention kernels at the bench batch, DRAM traffic of every kernel of one
# iteration at b = 2048 (roofline.traffic), launch list
mkdir -p gpurun_out
export BATCH_SCALE=2
python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_plain.log 2>&1 || { tail -5 gpurun_out/ncu_kernels_plain.log; exit 1; }
timeout -k 10 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"halo|igemm|attn|stem" -o /tmp/r02_kernels_full -f python scripts/ncu_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/ncu_summarize.py /tmp/r02_kernels_full.ncu-rep > gpurun_out/r02_ncu_full_kernels_b2048.txt
grep ran gpurun_out/ncu_kernels_plain.log >> gpurun_out/r02_ncu_full_kernels_b2048.txt
sz=$(stat -c %s /tmp/r02_kernels_full.ncu-rep); [ "$sz" -lt 50000000 ] && cp /tmp/r02_kernels_full.ncu-rep gpurun_out/
python scripts/iter_once.py --batch 2048 > gpurun_out/iter_plain.log 2>&1 &&
timeout -k 10 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_traffic_b2048.csv python scripts/iter_once.py --batch 2048 > gpurun_out/ncu.log 2>&1
python scripts/agg_traffic.py gpurun_out/r02_traffic_b2048.csv > gpurun_out/r02_traffic_b2048.json
head -c 1500 gpurun_out/r02_traffic_b2048.json
