#!/bin/bash
echo "== default (transposed drain)"; timeout 300 python tools/ce_time.py 128 64 2>&1 | grep "d="
echo "== direct 32B drain"; TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200_dd.so timeout 300 python tools/ce_time.py 128 64 2>&1 | grep "d="
TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200_dd.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/history_timeline.py 2>&1 | grep -v -i "warn\|return Variable" > gpurun_out/s40_history_timeline.txt
grep "total span" gpurun_out/s40_history_timeline.txt; grep -A14 "per kernel name" gpurun_out/s40_history_timeline.txt | cut -c1-110
timeout 300 python tools/step_timeline.py 2>&1 | grep "total span\|tower_fwd"
timeout 600 python -m pytest tests/test_gpu_history.py tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
