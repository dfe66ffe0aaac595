#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -rf -x -k "topk or rank" > gpurun_out/tk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tk_pytest.log
timeout 100 python scripts/time_topk.py 2>gpurun_out/tk.err | tee gpurun_out/tk_new2.jsonl | cut -c40-200
