#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
C=$PWD/two_tower_models_b200/csrc
echo "== umma rate"; timeout 120 tools/micro/umma_rate 2>&1 | tee $O/s5_umma_rate.txt
echo "== trace dU v3 CTA 2 (boundary)"; TT_CE_TRACE_CTA=2 TT_B200_LIB=$C/libtt_b200_bringup.so timeout 300 python tools/trace_ce.py 128 > $O/s5_trace_dU_cta2.txt 2>&1; tail -46 $O/s5_trace_dU_cta2.txt
