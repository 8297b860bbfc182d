#!/bin/bash
# Round-2 visit H (2 GPUs): in-segment sharding parity after the 8-way generalisation, 2-GPU bench line.
set -u
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -30 ) > gpurun_out/r2h_pytest_shard.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
tail -n 5 gpurun_out/r2h_pytest_shard.log; cut -c1-200 gpurun_out/r2h_bench_2gpu.json; tail -n 3 gpurun_out/r2h_bench_2gpu.err
