#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --sustain-seconds 0 > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_b4096.csv python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --sustain-seconds 0 > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/r2_launches_b4096.csv | head -60
