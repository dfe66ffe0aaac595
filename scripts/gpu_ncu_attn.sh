#!/bin/bash
mkdir -p gpurun_out
K=${1:-attention_umma_fwd}
OUT=${2:-r2_ncu_attn_umma_fwd}
timeout 600 ncu --set full --import-source on --clock-control none -f -k regex:$K -s 3 -c 1 -o gpurun_out/$OUT python scripts/time_attention.py 4096 52 2 32 > gpurun_out/$OUT.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/$OUT.log
