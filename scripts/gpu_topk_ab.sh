#!/bin/bash
# top-k over materialised rows: correctness subset, then the kernel variants side by side
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -rf -x -k "topk or rank or recall or ndcg or metric" > gpurun_out/tk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tk_pytest.log
echo "--- 256-thread sampled kernel (rows of 16K..128K scores), fixed-grid redo pass"
timeout 200 python scripts/time_topk.py 2>gpurun_out/tk.err | tee gpurun_out/tk_new.jsonl | cut -c40-200
echo "--- previous 512-thread sampled kernel (B4CP_TOPK_SMALL=0), fixed-grid redo pass"
B4CP_TOPK_SMALL=0 timeout 200 python scripts/time_topk.py 2>>gpurun_out/tk.err | tee gpurun_out/tk_old.jsonl | cut -c40-200
