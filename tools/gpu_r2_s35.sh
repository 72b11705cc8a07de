#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
timeout 300 python tools/step_timeline.py 2>&1 | grep -v -i "warn\|return Variable" | tail -22 | cut -c1-110
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s35_bench.json 2> $O/s35_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s35_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity']['loss_rel_vs_oracle'], 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'])
for k,v in d['configs'].items():
    print(k, 'value', round(v['value']), 'ms', round(v['ms_per_step'],4), 'roofline', v.get('roofline') and (v['roofline'].get('kernel'), round(v['roofline']['frac'],3)))
PY
