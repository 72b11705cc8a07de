#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mips.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for n in "" _t96 _t128 _t208; do
TT_B200_LIB=$PWD/two_tower_models_b200/csrc/libtt_b200$n.so timeout 600 python bench.py --workload mips --steps 3 --warmup 1 --no-cpu-baseline > $O/s34_mips$n.json 2> $O/s34_mips$n.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s34_mips$n.json'))
    k=d['roofline']['kernels']
    print('variant [$n]', 'value', round(d['value']), 'screen ms', round(k['mips_screen_kernel']['ms'],2), 'finalize', round(k['mips_finalize_kernel']['ms'],2), 'TF/s', round(d['roofline']['achieved']), d.get('parity'))
except Exception as ex: print('variant [$n] ERR', ex)
PY
done
