#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
echo "== fused zero fill"; timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6
echo "== side-stream memsets"; TT_B200_FUSED_ZERO_FILL=0 timeout 300 python tools/step_breakdown.py 2>&1 | grep -v -i "warn\|return Variable" | tail -6
timeout 300 python tools/step_timeline.py 2>&1 | grep -v -i "warn\|return Variable" | tail -22 | cut -c1-110
timeout 900 python bench.py --steps 20 --warmup 5 --no-extra-legs > $O/s29_bench.json 2> $O/s29_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s29_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity']['loss_rel_vs_oracle'])
PY
