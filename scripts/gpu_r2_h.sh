#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vocab_ce_gpu.py tests/test_kernels_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q --timeout 300 -rf -k "topk or score or filter or fullsize" 2>&1 | tail -5
timeout 300 python scripts/time_c5_topk.py 1000000 256 1 256 4096 16384 | sed 's/^/gen2 /'
B4CP_FILTER_GEN1=1 timeout 300 python scripts/time_c5_topk.py 1000000 256 4096 | sed 's/^/gen1 /'
timeout 300 python scripts/time_c5_topk.py 1000000 128 4096 | sed 's/^/gen2 h128 /'
