#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mips.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --workload mips --steps 3 --warmup 1 --no-cpu-baseline > $O/s30_mips128.json 2> $O/s30_mips128.err; echo rc=$?
python - <<'PY'
import json
for n in ('128',):
    try:
        d=json.load(open('gpurun_out/s30_mips%s.json'%n))
        print(n, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']), 'roofline', d.get('roofline'), 'parity', d.get('parity'))
    except Exception as ex:
        print(n, 'ERR', ex)
PY
tail -3 $O/s30_mips128.err
