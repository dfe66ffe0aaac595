#!/bin/bash
# LayerNorm backward prefetch: parity subset, kernel timing, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_zz_fit_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -rf -x > gpurun_out/ln_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ln_pytest.log
timeout 200 python scripts/time_ln.py > gpurun_out/ln_timing.jsonl 2> gpurun_out/ln_timing.err; echo "ln rc=$?"; cat gpurun_out/ln_timing.jsonl
bash scripts/gpu_bench_quick.sh
