#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ce_bwd3_kernel" -s 4 -c 2 -f -o gpurun_out/prof_ce_bwd3 python tools/ce_once.py 128 > gpurun_out/s10_ncu.log 2>&1
tail -3 gpurun_out/s10_ncu.log; ls -la gpurun_out/*.ncu-rep
