#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-extra-legs > $O/s37_bench.json 2> $O/s37_bench.err; echo "bench rc=$?"; tail -3 $O/s37_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s37_bench.json'))
r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'frac', round(r['frac'],3), 'raw', round(r['frac_raw_span'],3), 'ovh us', round(r['event_span_overhead_ms']*1e3,2), r['kernel'])
for k,v in r['kernels'].items(): print(' ', k, round(v['ms']*1e3,1), 'us x', v['launches_per_step'], ' eager', r['kernels_eager_ms'] and round(r['kernels_eager_ms'].get(k,0)*1e3,1))
PY
