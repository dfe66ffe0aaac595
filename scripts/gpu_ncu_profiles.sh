#!/bin/bash
# round-2 ncu captures of the final kernels (one GPU; each program exits 0 without ncu first)
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
python scripts/time_vocab.py 28672 128 54293 > gpurun_out/r2p_c1_plain.log 2>&1 || exit 1
timeout 600 $NCU -k regex:vocab_ce_fwd_ts -s 15 -c 1 -o gpurun_out/r2_ncu_c1_fwd python scripts/time_vocab.py 28672 128 54293 > gpurun_out/r2p_c1_fwd.log 2>&1; echo "c1 fwd rc=$?"
timeout 600 $NCU -k regex:vocab_ce_bwd_ts -s 4 -c 1 -o gpurun_out/r2_ncu_c1_bwd python scripts/time_vocab.py 28672 128 54293 > gpurun_out/r2p_c1_bwd.log 2>&1; echo "c1 bwd rc=$?"
timeout 600 $NCU -k regex:vocab_ce_dx -s 4 -c 1 -o gpurun_out/r2_ncu_c1_dx python scripts/time_vocab.py 28672 128 54293 > gpurun_out/r2p_c1_dx.log 2>&1; echo "c1 dx rc=$?"
python scripts/time_vocab.py 7424 256 1000000 > gpurun_out/r2p_c4_plain.log 2>&1 || exit 1
timeout 900 $NCU -k regex:vocab_ce_fwd_ts -s 15 -c 1 -o gpurun_out/r2_ncu_c4_fwd python scripts/time_vocab.py 7424 256 1000000 > gpurun_out/r2p_c4_fwd.log 2>&1; echo "c4 fwd rc=$?"
timeout 900 $NCU -k regex:vocab_ce_bwd_ts -s 4 -c 1 -o gpurun_out/r2_ncu_c4_bwd python scripts/time_vocab.py 7424 256 1000000 > gpurun_out/r2p_c4_bwd.log 2>&1; echo "c4 bwd rc=$?"
python scripts/time_attention.py 4096 52 2 32 > gpurun_out/r2p_attn_plain.log 2>&1 || exit 1
timeout 600 $NCU -k regex:attention_mma -s 6 -c 2 -o gpurun_out/r2_ncu_attn_c1 python scripts/time_attention.py 4096 52 2 32 > gpurun_out/r2p_attn.log 2>&1; echo "attn rc=$?"
timeout 600 $NCU -k regex:attention_mma_bwd -s 3 -c 1 -o gpurun_out/r2_ncu_attn_c1_bwd python scripts/time_attention.py 4096 52 2 32 > gpurun_out/r2p_attn_bwd.log 2>&1; echo "attn bwd rc=$?"
python scripts/one_c5.py 4096 256 > gpurun_out/r2p_c5_plain.log 2>&1 || exit 1
timeout 600 $NCU -k regex:score_filter -s 1 -c 1 -o gpurun_out/r2_ncu_c5_filter python scripts/one_c5.py 4096 256 > gpurun_out/r2p_c5.log 2>&1; echo "c5 rc=$?"
cat gpurun_out/r2p_c1_plain.log gpurun_out/r2p_c4_plain.log gpurun_out/r2p_attn_plain.log
ls -la gpurun_out/*.ncu-rep | tail -12
