mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "embed" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/check_data_parallel.py > gpurun_out/r2_data_parallel_2gpu.jsonl 2> gpurun_out/r2_dp_check.err; echo "dp check rc=$?"
tail -c 600 gpurun_out/r2_dp_check.err; cat gpurun_out/r2_data_parallel_2gpu.jsonl | cut -c1-1500
