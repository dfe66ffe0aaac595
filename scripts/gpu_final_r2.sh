#!/bin/bash
# final single-GPU evidence of round 2 (second session): smoke, full bench, reference arm, kernel timings
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2h_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/r2h_ref.json 2>gpurun_out/r2h_ref.err; echo "ref rc=$?"
timeout 600 python scripts/time_hbm_kernels.py > gpurun_out/r2h_hbm_kernels.jsonl 2> gpurun_out/r2h_hbm.err; echo "hbm rc=$?"; cut -c1-200 gpurun_out/r2h_hbm_kernels.jsonl
(timeout 300 python scripts/time_attention.py 4096 52 2 32; B4CP_ATTN_MMA_SYNC=1 timeout 300 python scripts/time_attention.py 4096 52 2 32; timeout 300 python scripts/time_attention.py 1024 103 4 32; timeout 300 python scripts/time_attention.py 256 202 4 64) > gpurun_out/r2h_attention.log 2>&1; cat gpurun_out/r2h_attention.log
(timeout 120 python scripts/time_gemm.py; timeout 120 python scripts/time_head_gemm.py) > gpurun_out/r2h_gemm.log 2>&1; cat gpurun_out/r2h_gemm.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'sessions',(d.get('e2e_from_sessions') or {}).get('value'),'b512',(d.get('b512') or {}).get('value'), 'sustained', (d.get('sustained') or {}).get('value'))
r=d['roofline']; print('fwd',r['ms_per_launch'],r['frac'],'bwd',r['other_kernels']['vocab_ce_bwd_ts_kernel']['ms_per_launch'],'stage',r['other_kernels']['vocab_stage_fwd_dx_bwd'])
print('topk', d.get('topk_queries_per_sec'), 'c4', json.dumps(d.get('c4_train'))[:300])
print('fp32', json.dumps(d.get('fp32_mode'))[:200])
PY
