#!/bin/bash
# ncu launch list of the benchmark step (eager launches, cold-cache serialised: compare shares) + full capture of CE kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
python tools/trace_ce.py 2>&1 | tail -30 | sed -n 12,15p
