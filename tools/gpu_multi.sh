#!/bin/bash
# usage: tools/gpu_multi.sh N   (N GPUs of one box): NCCL parity tests (kept in gpurun_out/), then bench.py on N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/test_gpu_distributed.txt 2>&1
echo "distributed tests rc=$?"; tail -5 gpurun_out/test_gpu_distributed.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -8
