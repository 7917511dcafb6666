#!/bin/bash
# round 2, call AE: ncu --set full of the conv / attention kernels on the hot shapes at the default bench batch (b = 4096)
mkdir -p gpurun_out
export BATCH_SCALE=4
python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_plain.log 2>&1 || { tail -5 gpurun_out/ncu_kernels_plain.log; exit 1; }
timeout -k 10 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"halo|igemm|attn|stem" -o /tmp/r02_kernels_full_b4096 -f python scripts/ncu_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/ncu_summarize.py /tmp/r02_kernels_full_b4096.ncu-rep > gpurun_out/r02_ncu_full_kernels_b4096.txt
grep ran gpurun_out/ncu_kernels_plain.log >> gpurun_out/r02_ncu_full_kernels_b4096.txt
grep -c "== launch" gpurun_out/r02_ncu_full_kernels_b4096.txt
grep "Kernel Name\|gpu__time_duration\|pipe_tensor\|dram__bytes_read\|lts__throughput" gpurun_out/r02_ncu_full_kernels_b4096.txt | head -60
