#!/bin/bash
# round 2, session 1: timelines of the CE backward (default build), LEAN build parity + timing, fused tower backward parity
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/s1_gpu.txt 2>&1
echo "== trace dU" ; timeout 300 python tools/trace_ce.py 128 > $O/s1_trace_dU.txt 2>&1; tail -3 $O/s1_trace_dU.txt | cut -c1-150
echo "== trace dV" ; timeout 300 python tools/trace_ce.py 128 dv > $O/s1_trace_dV.txt 2>&1; tail -3 $O/s1_trace_dV.txt | cut -c1-150
echo "== ce_time default"; timeout 300 python tools/ce_time.py 64 128 256 2>&1 | tee $O/s1_ce_time_default.txt | tail -3
echo "== ce_time lean"; TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200_lean.so timeout 300 python tools/ce_time.py 64 128 256 2>&1 | tee $O/s1_ce_time_lean.txt | tail -3
echo "== LEAN parity"
for f in test_gpu_kernels test_gpu_models; do
  TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200_lean.so timeout 900 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 600 > $O/s1_lean_$f.txt 2>&1
  echo "lean $f rc=$?"; tail -3 $O/s1_lean_$f.txt
done
echo "== fused tower bwd parity"
TT_B200_FUSED_TOWER_BWD=1 timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider --timeout 600 > $O/s1_ftb_models.txt 2>&1
echo "ftb rc=$?"; tail -5 $O/s1_ftb_models.txt
echo "== step breakdown default"; timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6 | tee $O/s1_breakdown_default.txt
echo "== step breakdown fused tower bwd"; TT_B200_FUSED_TOWER_BWD=1 timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6 | tee $O/s1_breakdown_ftb.txt
