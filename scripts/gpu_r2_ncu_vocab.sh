#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:vocab_ce_fwd_ts -s 15 -c 1 -o gpurun_out/r2_prof_fwd_dx -f python scripts/time_vocab.py 28672 128 54293 > gpurun_out/r2_ncu_fwd_dx.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_fwd_dx.log; ls -la gpurun_out/r2_prof_fwd_dx.ncu-rep
