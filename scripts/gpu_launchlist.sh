#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2g}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches_b4096.csv python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --sustain-seconds 0 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/${T}_launches_b4096.csv > gpurun_out/${T}_launch_summary_b4096.txt; head -40 gpurun_out/${T}_launch_summary_b4096.txt
