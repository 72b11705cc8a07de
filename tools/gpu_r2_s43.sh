#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/dp_timeline.py 2>&1 | grep " us \|total span" > gpurun_out/s43_dp_timeline_n$N.txt
tail -60 gpurun_out/s43_dp_timeline_n$N.txt | cut -c1-150
bash tools/gpu_scale.sh $N
