#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or dense or ln or colsum or adam" 2>&1 | tail -3
echo "two staging tiles:"; timeout 120 python scripts/time_gemm.py 2>&1 | tail -5
echo "one staging tile:"; B4CP_GEMM_ONE_STAGING=1 timeout 120 python scripts/time_gemm.py 2>&1 | tail -5
