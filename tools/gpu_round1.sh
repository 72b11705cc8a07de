#!/bin/bash
# first-light GPU run: parity tests, bench, ncu launch list + full capture of the CE kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x --deselect tests/test_gpu_models.py::test_requires_cuda_and_library > gpurun_out/pytest.txt 2>&1
tail -15 gpurun_out/pytest.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ce_ -s 12 -c 6 -f -o gpurun_out/prof_ce python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
