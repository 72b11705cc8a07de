#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -5
python tools/trace_ce.py 2>&1 | tail -30
python tools/bench_kernels.py small ce 2>&1 | tee gpurun_out/kbench_small.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_history.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -8
