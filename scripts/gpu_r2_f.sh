#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/time_vocab.py 28672 128 54293
timeout 120 python scripts/time_vocab.py 7424 256 1000000
timeout 600 python -m pytest tests/test_vocab_ce_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q --timeout 300 -rf -x 2>&1 | tail -3
