#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for dbg in 1 2; do
TT_B200_CHECK_IDS=0 TT_MIPS_DBG=$dbg timeout 600 python bench.py --workload mips --steps 2 --warmup 1 --no-cpu-baseline > $O/s32_mips$dbg.json 2> $O/s32_mips$dbg.err
python - <<PY
import json
d=json.load(open('gpurun_out/s32_mips$dbg.json'))
print('dbg', $dbg, 'screen ms', d['roofline']['kernels']['mips_screen_kernel']['ms'], 'TF/s', round(d['roofline']['achieved']))
PY
done
