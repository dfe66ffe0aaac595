#!/usr/bin/env bash
# One gpurun call that re-validates the whole tree on a fresh B200: GPU test suite, smoke, the
# N=1 bench line, the reference arm, and the small kernel timings whose numbers DESIGN.md quotes.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_full_check.sh'
# Everything is written under gpurun_out/full_check/ (merged back by gpurun).
set -u
out=gpurun_out/full_check
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$out/summary.txt"
tail -3 "$out/pytest_gpu.log" | tee -a "$out/summary.txt"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$out/summary.txt"
timeout 400 python bench.py --gpus 1 > "$out/bench_1gpu.json" 2> "$out/bench_1gpu.err"
echo "bench rc=$?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > "$out/bench_reference.json" 2>&1
echo "reference rc=$?" | tee -a "$out/summary.txt"
timeout 120 python scripts/time_cloze_build.py > "$out/cloze_build.jsonl" 2>&1
timeout 300 python scripts/time_hbm_kernels.py > "$out/hbm_kernels.jsonl" 2>&1
tail -c 600 "$out/bench_1gpu.json"
