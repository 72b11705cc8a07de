#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s17_bench.json 2> $O/s17_bench.err; echo "bench rc=$?"; tail -3 $O/s17_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s17_bench_ref.json 2> $O/s17_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s17_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'launches/step', d['launches_per_step'])
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), {k: round(v['ms']*1e3,1) for k,v in d['roofline']['kernels'].items()})
print('step', d['roofline'].get('step'))
print('parity', d.get('parity'))
print('cpu', d.get('cpu_baseline'))
for k,v in d.get('configs',{}).items():
    if v is None: print(k, None); continue
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    print(k, 'value', round(v['value']), v['unit'], 'ms', round(v['ms_per_step'],4), 'roofline', v.get('roofline') and round(v['roofline']['frac'],3), 'parity', v.get('parity'))
r=json.load(open('gpurun_out/s17_bench_ref.json')); print('ref', round(r['value']), r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
