#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"], "launches/step", d["launch_mode"]["kernels_per_step"])
print("roof", {k: d["roofline"][k] for k in ("frac", "ms_per_launch", "achieved", "kernel_share_of_step")})
print("b512", d["b512"]["value"], d["b512"]["ms_per_step"], "fp32", d["fp32_mode"]["value"], "topk", d["topk"]["value"])
print("c4", d["c4_train"]["value"], d["c4_train"]["ms_per_step"], d["c4_train"]["c5_topk"]["value"])
print("c5", d["c5_topk"]["value"], d["c5_topk"]["frac"], [(s["batch"], round(s["queries_per_sec"]), round(s["roofline"]["frac"],3)) for s in d["c5_topk"]["sweep"]], d["c5_topk"]["materialised_path"]["queries_per_sec"])
PY
