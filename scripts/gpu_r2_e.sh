#!/bin/bash
mkdir -p gpurun_out
for pp in 0 2 1; do
  B4CP_FWD_POLY=$pp timeout 120 python scripts/time_vocab.py 28672 128 54293 | sed "s/^/poly=$pp /"
done
for pp in 0 2 1; do
  B4CP_FWD_POLY=$pp timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed "s/^/poly=$pp /"
done
B4CP_FWD_POLY=1 timeout 600 python -m pytest tests/test_vocab_ce_gpu.py -m gpu -q --timeout 300 -rf -x 2>&1 | tail -3
B4CP_FWD_POLY=2 timeout 600 python -m pytest tests/test_vocab_ce_gpu.py -m gpu -q --timeout 300 -rf -x 2>&1 | tail -3
