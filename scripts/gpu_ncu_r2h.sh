#!/bin/bash
# ncu evidence of the final kernels: launch list of the step, full captures of the tcgen05 attention
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2h_launches_b4096.csv python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --no-cpu --sustain-seconds 0 > gpurun_out/r2h_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/r2h_launches_b4096.csv > gpurun_out/r2h_launch_summary_b4096.txt; head -14 gpurun_out/r2h_launch_summary_b4096.txt
for k in attention_umma_fwd attention_umma_bwd; do
  timeout 600 ncu --set full --import-source on --clock-control none -f -k regex:$k -s 3 -c 1 -o gpurun_out/r2h_ncu_$k python scripts/time_attention.py 4096 52 2 32 > gpurun_out/r2h_ncu_$k.log 2>&1; echo "$k rc=$?"
done
timeout 600 ncu --set full --import-source on --clock-control none -f -k regex:gemm_umma_persistent -s 3 -c 1 -o gpurun_out/r2h_ncu_gemm_qkv python scripts/time_gemm.py > gpurun_out/r2h_ncu_gemm_qkv.log 2>&1; echo "gemm rc=$?"
