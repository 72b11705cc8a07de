#!/bin/bash
# round-end evidence on one B200: parity tests, smoke, the three bench workloads, the reference arm, ncu launch list and
# full captures of the dominant kernels.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --workload history --steps 20 --warmup 5 > gpurun_out/bench_history.json 2> gpurun_out/bench_history.err; echo "history rc=$?"
timeout 900 python bench.py --workload mips --steps 2 --warmup 1 > gpurun_out/bench_mips.json 2> gpurun_out/bench_mips.err; echo "mips rc=$?"
python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6 > gpurun_out/step_breakdown.txt; cat gpurun_out/step_breakdown.txt
python tools/step_timeline.py 2>&1 | grep -v -i "warn\|return Variable" | tail -24 | cut -c1-110 > gpurun_out/step_timeline.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ce_fwd_kernel|ce_bwd2_kernel|tower_fwd_kernel|adam_kernel|ce_bwd_reduce|ce_combine_loss" -s 12 -c 8 -f -o gpurun_out/prof_step_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv 2>&1 | tail -4
python -c "
import json
for f in ['bench_n1','bench_history','bench_mips','bench_ref']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value']), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
"
