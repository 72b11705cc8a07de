#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "inbatch" > $O/s26_kernels.txt 2>&1; echo "rc=$?"; tail -8 $O/s26_kernels.txt
echo "== v3x"; timeout 300 python tools/ce_time.py 256 2>&1 | tail -1
echo "== v2 "; TT_CE_BWD_V3X=0 timeout 300 python tools/ce_time.py 256 2>&1 | tail -1
