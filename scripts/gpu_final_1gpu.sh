#!/bin/bash
# final single-GPU measurements of round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/r2f_ref.json 2>gpurun_out/r2f_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 600 python scripts/time_hbm_kernels.py > gpurun_out/r2f_hbm_kernels.jsonl 2> gpurun_out/r2f_hbm.err; echo "hbm rc=$?"; cat gpurun_out/r2f_hbm_kernels.jsonl | cut -c1-260
timeout 300 python scripts/time_attention.py 4096 52 2 32 > gpurun_out/r2f_attention.log 2>&1; timeout 300 python scripts/time_attention.py 256 202 4 64 >> gpurun_out/r2f_attention.log 2>&1; cat gpurun_out/r2f_attention.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_launches_b4096.csv python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-fp32 --sustain-seconds 0 > gpurun_out/r2f_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/r2f_launches_b4096.csv > gpurun_out/r2f_launch_summary_b4096.txt; head -12 gpurun_out/r2f_launch_summary_b4096.txt
