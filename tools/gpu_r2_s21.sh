#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_history.py tests/test_gpu_debias.py -m gpu -q -p no:cacheprovider --timeout 600 -x > $O/s21_history.txt 2>&1; echo "rc=$?"; tail -15 $O/s21_history.txt
timeout 900 python bench.py --workload history --steps 20 --warmup 5 > $O/s21_bench_history.json 2> $O/s21_bench_history.err; echo "bench rc=$?"; tail -3 $O/s21_bench_history.err
python - <<'PY'
import json
h=json.load(open('gpurun_out/s21_bench_history.json'))
print(h['value'], h['ms_per_step'], h['launches_per_step'], h['loss'])
for k,v in sorted(h['kernel_ms_per_step'].items(), key=lambda kv:-kv[1]): print(f"{k:40s} {v*1e3:8.1f} us")
PY
