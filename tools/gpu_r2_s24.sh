#!/bin/bash
C=$PWD/two_tower_models_b200/csrc
for i in 1 2; do
echo "== default"; timeout 300 python tools/ce_time.py 64 128 256 2>&1 | tail -3
echo "== fwd poly"; TT_B200_LIB=$C/libtt_b200_fpoly.so timeout 300 python tools/ce_time.py 64 128 256 2>&1 | tail -3
done
TT_B200_LIB=$C/libtt_b200_fpoly.so timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x -k "inbatch_ce_forward or full_size" 2>&1 | tail -3
