#!/bin/bash
# 2-GPU: data-parallel correctness against one process on the concatenated batch, the vocabulary-
# parallel layer against the replicated one, and the bench line at N = 2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/check_data_parallel.py > gpurun_out/r2h_data_parallel_2gpu.jsonl 2> gpurun_out/r2h_dp_check.err; echo "dp check rc=$?"
tail -c 400 gpurun_out/r2h_dp_check.err; cut -c1-900 gpurun_out/r2h_data_parallel_2gpu.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/check_vocab_parallel.py > gpurun_out/r2h_vocab_parallel_2gpu.jsonl 2> gpurun_out/r2h_vp_check.err; echo "vp check rc=$?"
tail -c 400 gpurun_out/r2h_vp_check.err; cut -c1-900 gpurun_out/r2h_vocab_parallel_2gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-fp32 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench N=2 rc=$?"
tail -c 600 gpurun_out/r2h_bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2h_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"])
print("b512", d["b512"]["value"], "topk", d["topk"]["value"], "c4", d["c4_train"]["value"], d["c4_train"]["ms_per_step"])
PY
