#!/bin/bash
# usage: tools/gpu_multi.sh N   (N GPUs of one box)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider --timeout 200 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; cat gpurun_out/bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --peer-ce --no-cpu-baseline > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err
echo "peer bench rc=$?"; cat gpurun_out/bench_n${N}_peer.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_peer.err | tail -8
