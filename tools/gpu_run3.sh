#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py gemm ce attn > gpurun_out/kbench.txt 2>&1; cat gpurun_out/kbench.txt
timeout 900 python tools/bench_kernels.py mips > gpurun_out/kbench_mips.txt 2>&1; cat gpurun_out/kbench_mips.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_eager.json
bash tools/gpu_tests.sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
