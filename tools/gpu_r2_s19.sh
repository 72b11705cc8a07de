#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/history_kernels_time.py 2>&1 | grep -v -i "warn" | tee gpurun_out/s19_history_kernels.txt
