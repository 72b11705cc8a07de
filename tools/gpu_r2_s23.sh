#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 600 -x -k "inbatch" > $O/s23_kernels.txt 2>&1; echo "rc=$?"; tail -3 $O/s23_kernels.txt
for g in 0 2 4 6 8; do echo "== ghost $g"; TT_CE_BWD_GHOST=$g timeout 300 python tools/ce_time.py 128 2>&1 | tail -1; done
timeout 900 python bench.py --steps 20 --warmup 5 --no-extra-legs > $O/s23_bench.json 2> $O/s23_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s23_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']))
PY
