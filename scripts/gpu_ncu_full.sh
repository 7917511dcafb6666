#!/bin/bash
# ncu --set full capture of the conv-engine / attention kernels on the hot shapes (after a plain run exits 0).
# The per-launch summary is produced on the box; the .ncu-rep travels back only if it fits gpurun's 64 MiB limit.
mkdir -p gpurun_out
python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_plain.log 2>&1 || { tail -5 gpurun_out/ncu_kernels_plain.log; exit 1; }
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"${NCU_REGEX:-halo|igemm|attn|stem}" -o /tmp/r01d_kernels_full -f python scripts/ncu_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la /tmp/*.ncu-rep
python scripts/ncu_summarize.py /tmp/r01d_kernels_full.ncu-rep > gpurun_out/r01d_ncu_full_summary.txt
grep ran gpurun_out/ncu_kernels_plain.log >> gpurun_out/r01d_ncu_full_summary.txt
sz=$(stat -c %s /tmp/r01d_kernels_full.ncu-rep); [ "$sz" -lt 50000000 ] && cp /tmp/r01d_kernels_full.ncu-rep gpurun_out/
wc -l gpurun_out/r01d_ncu_full_summary.txt
