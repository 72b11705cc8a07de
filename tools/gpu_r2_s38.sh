#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/history_timeline.py 2>&1 | grep -v -i "warn\|return Variable" > gpurun_out/s38_history_timeline.txt
tail -45 gpurun_out/s38_history_timeline.txt | cut -c1-130
