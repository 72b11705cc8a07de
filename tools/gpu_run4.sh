#!/bin/bash
mkdir -p gpurun_out
python tools/trace_gemm.py fwd 2>&1 | tail -8
python tools/trace_gemm.py splitk 2>&1 | tail -8
python tools/bench_kernels.py small 2>&1 | tee gpurun_out/kbench_small.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
bash tools/gpu_tests.sh
