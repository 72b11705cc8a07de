#!/bin/bash
# round 2, session 2: first light of the v3 CE backward (ce_bwd3.cu)
mkdir -p gpurun_out
O=gpurun_out
C=$PWD/two_tower_models_b200/csrc
echo "== kernels parity (v3 default)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x > $O/s12_kernels.txt 2>&1; echo "rc=$?"; tail -15 $O/s12_kernels.txt
echo "== ce_time v3"; timeout 300 python tools/ce_time.py 64 128 2>&1 | tee $O/s12_ce_time_v3.txt | tail -3
echo "== ce_time v3 poly"; TT_B200_LIB=$C/libtt_b200_poly.so timeout 300 python tools/ce_time.py 64 128 2>&1 | tee $O/s12_ce_time_poly.txt | tail -3
echo "== ce_time v2"; TT_CE_BWD_V3=0 timeout 300 python tools/ce_time.py 128 2>&1 | tee $O/s12_ce_time_v2.txt | tail -3
echo "== trace dU v3"; TT_B200_LIB=$C/libtt_b200_bringup.so timeout 300 python tools/trace_ce.py 128 > $O/s12_trace_dU.txt 2>&1; head -8 $O/s12_trace_dU.txt | cut -c1-200; tail -4 $O/s12_trace_dU.txt
echo "== trace dV v3"; TT_B200_LIB=$C/libtt_b200_bringup.so timeout 300 python tools/trace_ce.py 128 dv > $O/s12_trace_dV.txt 2>&1; head -8 $O/s12_trace_dV.txt | cut -c1-200; tail -4 $O/s12_trace_dV.txt
echo "== models parity"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_debias.py -m gpu -q -p no:cacheprovider --timeout 600 > $O/s12_models.txt 2>&1; echo "rc=$?"; tail -5 $O/s12_models.txt
echo "== step breakdown"; timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6 | tee $O/s12_breakdown.txt
