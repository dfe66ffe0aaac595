#!/bin/bash
# Focused GPU run of the vocabulary-stage tests, one pytest process per group so that a hung
# kernel (killed by the timeout) does not hide the other groups' results.
mkdir -p gpurun_out
for grp in "test_fused_forward_lse_and_target" "test_fused_lazy_rescale" "test_fused_backward_gradients" "test_fused_all_rows_padded or test_vocab_shards_merge"; do
  echo "=== $grp" >> gpurun_out/vocab_tests.log
  timeout 240 python -m pytest tests/test_vocab_ce_gpu.py -q -k "$grp" --timeout 120 --timeout-method thread 2>&1 | tail -40 >> gpurun_out/vocab_tests.log
  echo "rc=$?" >> gpurun_out/vocab_tests.log
done
tail -60 gpurun_out/vocab_tests.log
