#!/bin/bash
# LayerNorm backward: prefetching loop (tree) against the previous kernel.  Build the other side first:
#   git show <rev>:bert4clickpath_b200/csrc/encoder.cu > /tmp/e.cu && nvcc <build.py FLAGS> -c /tmp/e.cu -o /tmp/e.o
#   nvcc -shared -cudart shared -o scratch_ab/libb4cp_old.so $(ls bert4clickpath_b200/csrc/build/*.o | grep -v /encoder.o) /tmp/e.o
# (scratch_ab/ is not tracked; *.so files travel to the GPU box with the tree)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_fit_gpu.py -m gpu -q --timeout 300 -rf -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
for i in 1 2; do
  echo "--- new"; timeout 200 python scripts/time_ln.py 2>&1 | grep bwd | cut -c1-60,150-260
  echo "--- old"; timeout 200 python scripts/with_lib.py scratch_ab/libb4cp_old.so scripts/time_ln.py 2>&1 | grep bwd | cut -c1-60,150-260
done
B="bench.py --steps 40 --warmup 10 --no-c4 --no-topk --no-fp32 --sustain-seconds 0 --no-cpu --no-builder"
for i in 1 2; do
  timeout 300 python $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['ms_per_step'], (d.get('b512') or {}).get('ms_per_step'))"
  timeout 300 python scripts/with_lib.py scratch_ab/libb4cp_old.so $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', d['ms_per_step'], (d.get('b512') or {}).get('ms_per_step'))"
done
