#!/bin/bash
mkdir -p gpurun_out
N=8
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --no-c4 --no-topk --no-builder --sustain-seconds 1.0 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms/step %.4f" % d["ms_per_step"], "sustained %.4f" % d["sustained"]["ms_per_step"], "b512 %.4f" % d["b512"]["ms_per_step"])
except Exception as e:
    print("$tag ERR", e)
PY
}
run default X=1
run nooverlap B4CP_DP_OVERLAP=0
run ch8 NCCL_MAX_NCHANNELS=8
run ch4 NCCL_MAX_NCHANNELS=4
run ch2 NCCL_MAX_NCHANNELS=2
run nvls0 NCCL_NVLS_ENABLE=0
run ctas8 NCCL_MAX_CTAS=8
