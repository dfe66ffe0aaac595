#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -f -k regex:vocab_ce_fwd_ts -s 8 -c 1 -o gpurun_out/r2_ncu_c1_fwd_instep python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --sustain-seconds 0 > gpurun_out/r2_ncu_instep.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2_ncu_c1_fwd_instep.ncu-rep
