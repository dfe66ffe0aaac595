#!/bin/bash
for v in 1 0; do
  B4CP_WEIGHT_STREAM=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-topk --no-b512 --no-fp32 --no-builder --no-cpu --sustain-seconds 0 > gpurun_out/c4_ab_$v.json 2> gpurun_out/c4_ab_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/c4_ab_$v.json'))
print('ws=$v', 'c1', d['ms_per_step'], 'c4', d['c4_train']['ms_per_step'], 'c5', d['c4_train']['c5_topk']['ms_per_step'])
PY
done
