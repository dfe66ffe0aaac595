#!/bin/bash
# full GPU test suite + a short bench (C1 legs only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-c4 --no-fp32 --sustain-seconds 0 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'sessions',(d.get('e2e_from_sessions') or {}).get('value'),'b512',(d.get('b512') or {}).get('value'))
print('strings', json.dumps(d.get('e2e_strings'))[:400])
print('topk', d.get('topk_queries_per_sec'), json.dumps(d.get('topk'))[:300])
r=d['roofline']; print('fwd',r['ms_per_launch'],r['frac'],'bwd',r['other_kernels']['vocab_ce_bwd_ts_kernel']['ms_per_launch'],'stage',r['other_kernels']['vocab_stage_fwd_dx_bwd']['ms_per_launch'],'launches',d['launch_mode']['kernels_per_step'])
PY
