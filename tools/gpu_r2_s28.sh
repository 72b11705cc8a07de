#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
TT_CE_BWD_X128=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "inbatch" > $O/s28_kernels.txt 2>&1; echo "rc=$?"; tail -4 $O/s28_kernels.txt
for i in 1 2; do
echo "== 96-col tiles, X in TMEM (default)"; timeout 300 python tools/ce_time.py 128 2>&1 | tail -1
echo "== 128-col tiles, X in smem"; TT_CE_BWD_X128=1 timeout 300 python tools/ce_time.py 128 2>&1 | tail -1
done
