#!/bin/bash
for sg in 0 1 3 9 27; do
B4CP_FWD_STAGGER=$sg timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed "s/^/stagger=$sg /" | cut -c1-150
done
B4CP_FWD_STAGGER=3 B4CP_FWD_NO_KSPLIT=1 timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed "s/^/stagger=3 whole /" | cut -c1-150
timeout 300 python -m pytest tests/test_vocab_ce_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
