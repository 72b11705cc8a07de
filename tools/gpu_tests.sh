#!/bin/bash
# parity tests, one pytest process per file so that a device-side trap in one file cannot poison the next
mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_models test_gpu_mips test_gpu_history test_gpu_debias test_gpu_optim; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 600 "$@" > gpurun_out/$f.txt 2>&1
  echo "== $f rc=$?"; tail -4 gpurun_out/$f.txt
done
