#!/bin/bash
# attention: parity tests on both paths + timing (each under its own timeout)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k attention > gpurun_out/attn_pytest.log 2>&1; echo "pytest umma rc=$?"; tail -15 gpurun_out/attn_pytest.log
for shape in "4096 52 2 32" "1024 103 4 32"; do
  timeout 120 python scripts/time_attention.py $shape 2>&1 | tail -3
  B4CP_ATTN_MMA_SYNC=1 timeout 120 python scripts/time_attention.py $shape 2>&1 | tail -3
done
