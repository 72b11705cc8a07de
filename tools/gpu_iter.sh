#!/bin/bash
# one optimisation iteration: parity tests, then the step breakdown and kernel micro-benchmarks
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed|Error|error" | head -20
python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -8
python tools/bench_kernels.py small ce 2>&1 | tail -18
