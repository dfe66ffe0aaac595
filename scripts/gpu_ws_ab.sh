#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log
for v in 1; do echo "weight stream=$v"; B4CP_WEIGHT_STREAM=$v bash scripts/gpu_bench_quick.sh 2>&1 | tail -2; done
