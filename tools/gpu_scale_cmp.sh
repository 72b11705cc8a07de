#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for c in 0 16 32; do
TT_B200_RS_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((c%10)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra-legs > gpurun_out/bench_n${N}_c$c.json 2> gpurun_out/bench_n${N}_c$c.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_c$c.json') if l.startswith('{')][-1])
r=d['roofline']
print('RS_CTAS=$c N=$N value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'dU', round(r['kernels']['ce_bwd3_kernel_dU']['ms']*1e3,1), 'dV', round(r['kernels']['ce_bwd3_kernel_dV']['ms']*1e3,1), 'kernels', round(r['device_ms_all_kernels'],4))
PY
done
