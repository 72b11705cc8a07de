#!/bin/bash
mkdir -p gpurun_out
python tools/step_breakdown.py 2>&1 | grep -v Warning | tail -8
python tools/bench_kernels.py small ce 2>&1 | tail -20
python tools/trace_ce.py 128 fwd 2>&1 | tail -30
python tools/trace_ce.py 128 2>&1 | tail -40
