#!/bin/bash
bash tools/gpu_r2_s30.sh 2>&1 | cut -c1-420
TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200_bringup.so timeout 600 python tools/mips_trace.py 2>&1 | grep -v -i "warn" | tail -16
