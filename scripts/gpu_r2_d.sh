#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vocab_ce_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q --timeout 300 -rf -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest(ranges) rc=$?"
tail -4 gpurun_out/r2d_pytest.log
B4CP_FWD_SCHED=grid timeout 600 python -m pytest tests/test_vocab_ce_gpu.py -m gpu -q --timeout 300 -rf -x > gpurun_out/r2d_pytest_grid.log 2>&1; echo "pytest(grid) rc=$?"
tail -3 gpurun_out/r2d_pytest_grid.log
for M in 28672 3584; do
  timeout 120 python scripts/time_vocab.py $M 128 54293 | sed 's/^/ranges /'
done
timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed 's/^/grid   /'
timeout 120 python scripts/time_vocab.py 7424 256 54293 | sed 's/^/ranges h256 /'
