#!/bin/bash
mkdir -p gpurun_out
echo "== umma rate (unrolled)"; timeout 120 tools/micro/umma_rate 2>&1 | tee gpurun_out/s6_umma_rate.txt
