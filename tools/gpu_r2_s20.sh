#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_fwd_tc|attn_bwd_tc" -s 4 -c 4 -f -o gpurun_out/prof_history_big python tools/history_kernels_once.py > gpurun_out/s20_ncu.log 2>&1
tail -2 gpurun_out/s20_ncu.log
