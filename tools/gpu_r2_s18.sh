#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/step_timeline.py 2>&1 | grep -v -i "warn\|return Variable" | tail -30 | cut -c1-130 | tee $O/s18_step_timeline.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-extra-legs > $O/s18_bench.json 2> $O/s18_bench.err; echo "bench rc=$?"; tail -3 $O/s18_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s18_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'launches/step', d['launches_per_step'])
PY
