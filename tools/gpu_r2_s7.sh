#!/bin/bash
mkdir -p gpurun_out
C=$PWD/two_tower_models_b200/csrc
TT_B200_LIB=$C/libtt_b200_bringup.so timeout 300 python tools/trace_ce.py 128 > gpurun_out/s7_trace_dU.txt 2>&1; sed -n 8,40p gpurun_out/s7_trace_dU.txt
