#!/bin/bash
# final single-GPU evidence of round 2 (third session): smoke, reference arm, full bench, kernel
# timings, ncu launch list of the step, full ncu captures of the kernels changed in this session
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/r2c_ref.json 2>gpurun_out/r2c_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_bench.err
timeout 600 python scripts/time_hbm_kernels.py > gpurun_out/r2c_hbm_kernels.jsonl 2> gpurun_out/r2c_hbm.err; echo "hbm rc=$?"; cut -c1-200 gpurun_out/r2c_hbm_kernels.jsonl
timeout 200 python scripts/time_topk.py > gpurun_out/r2c_topk.jsonl 2>gpurun_out/r2c_topk.err; cut -c1-200 gpurun_out/r2c_topk.jsonl
timeout 200 python scripts/time_ln.py > gpurun_out/r2c_ln.jsonl 2>gpurun_out/r2c_ln.err; cut -c1-200 gpurun_out/r2c_ln.jsonl
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'strings',(d.get('e2e_strings') or {}).get('value'),'sessions',(d.get('e2e_from_sessions') or {}).get('value'),'b512',(d.get('b512') or {}).get('value'), 'sustained', (d.get('sustained') or {}).get('value'))
r=d['roofline']; print('fwd',r['ms_per_launch'],r['frac'],'bwd',r['other_kernels']['vocab_ce_bwd_ts_kernel']['ms_per_launch'],'stage',r['other_kernels']['vocab_stage_fwd_dx_bwd'])
print('topk', d.get('topk_queries_per_sec'), 'c4', json.dumps(d.get('c4_train'))[:300])
print('fp32', json.dumps(d.get('fp32_mode'))[:200])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2c_launches_b4096.csv python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-strings --no-fp32 --no-cpu --sustain-seconds 0 > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/r2c_launches_b4096.csv > gpurun_out/r2c_launch_summary_b4096.txt; head -14 gpurun_out/r2c_launch_summary_b4096.txt
timeout 300 ncu --set full --import-source on --clock-control none -f -k regex:topk_sample_small -s 2 -c 1 -o gpurun_out/r2c_ncu_topk_small python scripts/time_topk.py > gpurun_out/r2c_ncu_topk_small.log 2>&1; echo "ncu topk rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -f -k regex:residual_ln_bwd -s 2 -c 1 -o gpurun_out/r2c_ncu_ln_bwd python scripts/time_ln.py > gpurun_out/r2c_ncu_ln_bwd.log 2>&1; echo "ncu ln rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -f -k regex:segment_tile_sum -s 4 -c 1 -o gpurun_out/r2c_ncu_segment_tile_sum python bench.py --steps 3 --warmup 3 --no-graph --no-c4 --no-topk --no-b512 --no-builder --no-strings --no-fp32 --no-cpu --sustain-seconds 0 > gpurun_out/r2c_ncu_seg.log 2>&1; echo "ncu seg rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
