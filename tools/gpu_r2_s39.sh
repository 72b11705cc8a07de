#!/bin/bash
timeout 600 python tools/history_kernels_time.py 2>&1 | grep -v -i "warn" | head -8
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_history.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
