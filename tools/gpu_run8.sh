#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-300
timeout 900 python bench.py --workload history --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_history.json 2> gpurun_out/bench_history.err; cat gpurun_out/bench_history.json | cut -c1-1500; tail -2 gpurun_out/bench_history.err | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
