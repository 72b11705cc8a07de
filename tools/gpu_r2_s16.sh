#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
C=$PWD/two_tower_models_b200/csrc
echo "== kernels parity (default, CW=32)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "inbatch" > $O/s16_kernels.txt 2>&1; echo "rc=$?"; tail -3 $O/s16_kernels.txt
echo "== kernels parity (CW=16)"
TT_B200_LIB=$C/libtt_b200_cw16.so timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "inbatch" > $O/s16_kernels_cw16.txt 2>&1; echo "rc=$?"; tail -3 $O/s16_kernels_cw16.txt
for i in 1 2; do echo "== ce_time CW=32"; timeout 300 python tools/ce_time.py 64 128 2>&1 | tee -a $O/s16_ce_time.txt | tail -2; 
echo "== ce_time CW=16"; TT_B200_LIB=$C/libtt_b200_cw16.so timeout 300 python tools/ce_time.py 64 128 2>&1 | tee -a $O/s16_ce_time_cw16.txt | tail -2; done
