#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_history.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -4
python tools/bench_kernels.py attn 2>&1 | tail -3
timeout 900 python bench.py --workload history --steps 10 --warmup 3 > gpurun_out/bench_history.json 2> gpurun_out/bench_history.err; cat gpurun_out/bench_history.json; tail -3 gpurun_out/bench_history.err
timeout 900 python bench.py --workload mips --steps 2 --warmup 1 > gpurun_out/bench_mips.json 2> gpurun_out/bench_mips.err; cat gpurun_out/bench_mips.json; tail -3 gpurun_out/bench_mips.err
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-400
