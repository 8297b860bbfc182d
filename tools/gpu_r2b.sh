#!/bin/bash
# Round-2 visit B (2 GPUs): in-segment sharding parity test, then the 2-GPU bench line (replicas + in-segment latency).
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -30 ) > gpurun_out/r2b_pytest_shard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err
tail -5 gpurun_out/r2b_pytest_shard.log
cut -c1-400 gpurun_out/r2b_bench_2gpu.json
tail -5 gpurun_out/r2b_bench_2gpu.err
