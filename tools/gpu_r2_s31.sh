#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mips_screen -c 1 -o $O/s31_mips -f python bench.py --workload mips --queries 37888 --steps 1 --warmup 0 --no-cpu-baseline > $O/s31_ncu.log 2>&1
echo rc=$?; tail -2 $O/s31_ncu.log | cut -c1-200; ls -la $O/s31_mips.ncu-rep
