#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -f -k regex:gemm_umma_persistent -s 3 -c 1 -o gpurun_out/r2_ncu_gemm_qkv python scripts/time_gemm.py > gpurun_out/r2_ncu_gemm_qkv.log 2>&1; echo "rc=$?"
