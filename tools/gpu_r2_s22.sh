#!/bin/bash
mkdir -p gpurun_out
echo "== late fill (default)"; timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6
echo "== early fill"; TT_B200_LATE_ZERO_FILL=0 timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6
