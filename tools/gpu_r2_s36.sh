#!/bin/bash
bash tools/gpu_tests.sh 2>&1 | grep -E "rc=|passed|failed"
