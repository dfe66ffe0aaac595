#!/bin/bash
# round-2 GPU call B (2 GPUs): GPU tests, smoke, bench at N=1 (short) and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
tail -3 gpurun_out/r2b_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-c4 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "bench2 rc=$?"
tail -c 2000 gpurun_out/r2b_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench1.json", "gpurun_out/r2b_bench2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "b512", d.get("b512"), "fp32", d.get("fp32_mode"), "c4", d.get("c4_train"))
        print("roof", d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
    except Exception as e:
        print(f, "ERR", e)
PY
