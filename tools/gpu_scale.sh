#!/bin/bash
# usage: tools/gpu_scale.sh N : bench.py on N GPUs the way the driver launches it (no tests)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$\|UserWarning\|run_backward\|AccumulateGrad" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
r=d['roofline']
print('N=$N value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']))
print({k: (round(v['ms']*1e3,1), v['launches_per_step']) for k,v in r['kernels'].items()})
print('device_ms_all_kernels', round(r['device_ms_all_kernels'],4))
for k,v in d['configs'].items(): print(k, v.get('error') or (round(v['value']), round(v['ms_per_step'],4), v['roofline']['kernel'], round(v['roofline']['frac'],3), {kk: round(vv['ms']*1e3,1) for kk,vv in v['roofline']['kernels'].items()}))
PY
