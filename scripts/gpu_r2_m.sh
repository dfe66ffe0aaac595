#!/bin/bash
timeout 120 python scripts/time_attention.py 4096 52 2 32
timeout 120 python scripts/time_attention.py 512 52 2 32
timeout 120 python scripts/time_attention.py 256 103 4 32
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3
