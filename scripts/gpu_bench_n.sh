#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"], d["sustained"]["ms_per_step"])
print("b512", d["b512"])
print("topk", d["topk"]["value"])
print("c4", d["c4_train"])
PY
