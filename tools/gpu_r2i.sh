#!/bin/bash
# Round-2 visit I (8 GPUs): in-segment sharding parity on 2 / 4 / 8 GPUs, then the 8-GPU bench line (replicas + one segment on 8 GPUs).
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2i_gpus.txt; nproc >> gpurun_out/r2i_gpus.txt
( timeout 400 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -30 ) > gpurun_out/r2i_pytest_shard.log
EXTRA=""
grep -q "3 passed" gpurun_out/r2i_pytest_shard.log || EXTRA="--no-in-segment"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 8 --steps 5 --warmup 3 $EXTRA > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench_8gpu.err
tail -n 6 gpurun_out/r2i_pytest_shard.log; cut -c1-200 gpurun_out/r2i_bench_8gpu.json; tail -n 3 gpurun_out/r2i_bench_8gpu.err
