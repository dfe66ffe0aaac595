#!/bin/bash
# round-2 GPU call A: environment probe, full GPU test suite, smoke, one bench line
mkdir -p gpurun_out
{ echo "PYTHONPATH=$PYTHONPATH"; python - <<'PY'
import sys
try:
    import sitecustomize
    print("sitecustomize:", sitecustomize.__file__)
    print(open(sitecustomize.__file__).read()[:6000])
except Exception as e:
    print("no sitecustomize:", e)
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc; } > gpurun_out/r2a_env.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
tail -3 gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
head -c 3000 gpurun_out/r2a_bench.json
