#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload history --steps 20 --warmup 5 > gpurun_out/bench_history.json 2> gpurun_out/bench_history.err; cat gpurun_out/bench_history.json; tail -2 gpurun_out/bench_history.err | cut -c1-200
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-200
# ncu: full-metric capture of the scoring kernels (eager path, few steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ce_fwd_kernel|ce_bwd2_kernel" -s 4 -c 3 -f -o gpurun_out/prof_ce_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_ce.log 2>&1
tail -2 gpurun_out/ncu_ce.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_" -s 0 -c 4 -f -o gpurun_out/prof_attn_r01 python bench.py --workload history --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
