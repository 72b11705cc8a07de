#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_history.py tests/test_gpu_models.py tests/test_gpu_debias.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
