#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6
timeout 300 python tools/step_timeline.py 2>&1 | grep -v -i "warn\|return Variable" | tail -12 | cut -c1-110
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
