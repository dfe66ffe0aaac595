#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_zz_reference_golden_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 200 -rf -x -k "embed or reference or golden or segment" > gpurun_out/eb_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/eb_pytest.log
timeout 100 python scripts/time_embed_bwd.py 2>gpurun_out/eb.err | tee gpurun_out/eb_new.jsonl | cut -c1-260; tail -2 gpurun_out/eb.err
