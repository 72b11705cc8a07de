#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
timeout 300 python tools/tower_time.py 2>&1 | grep -v -i warn | tail -40 | tee gpurun_out/s27_tower_time.txt
