#!/bin/bash
# ncu --set full capture of the conv-engine kernels on the hot shapes (after a plain run exits 0).
mkdir -p gpurun_out
python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_plain.log 2>&1 || { tail -5 gpurun_out/ncu_kernels_plain.log; exit 1; }
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"halo|igemm" \
    -o gpurun_out/r01_conv_kernels_full -f python scripts/ncu_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
