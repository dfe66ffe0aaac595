#!/bin/bash
mkdir -p gpurun_out
B4CP_FWD_OPT=1 timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed 's/^/opt=1 /'
B4CP_FWD_OPT=0 timeout 120 python scripts/time_vocab.py 7424 256 1000000 | sed 's/^/opt=0 /'
B4CP_FWD_OPT=0 timeout 120 python scripts/time_vocab.py 28672 128 54293 | sed 's/^/opt=0 /'
timeout 120 python scripts/time_vocab.py 28672 128 54293 | sed 's/^/default /'
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"])
print("roof", {k: d["roofline"][k] for k in ("frac", "ms_per_launch", "achieved", "kernel_share_of_step")})
print("bwd", d["roofline"]["other_kernels"]["vocab_ce_bwd_ts_kernel"]["frac"], "stage", d["roofline"]["other_kernels"]["vocab_stage_fwd_dx_bwd"]["frac"])
print("c4", d["c4_train"]["value"], d["c4_train"]["ms_per_step"], "c4stage", d["c4_vocab_stage"]["stage_fwd_dx_bwd"], d["c4_vocab_stage"]["vocab_ce_fwd_ts_kernel"])
PY
