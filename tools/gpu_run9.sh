#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh
python tools/bench_kernels.py small gemm 2>&1 | grep -E "gemm" | head -14
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('base ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'])"
timeout 900 python bench.py --workload history --steps 20 --warmup 5 > gpurun_out/bench_history.json 2> gpurun_out/bench_history.err
python -c "
import json; d=json.load(open('gpurun_out/bench_history.json')); print('history ms/step', d['ms_per_step'], 'value', d['value']); print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
# ncu evidence: launch list of one eager step + full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ce_fwd_kernel|ce_bwd2_kernel|gemm_kernel" -s 20 -c 8 -f -o gpurun_out/prof_step_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mips_" -c 2 -f -o gpurun_out/prof_mips_r01 python bench.py --workload mips --queries 16384 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01.csv
