#!/bin/bash
# Round-2 visit J (4 GPUs): the 4-GPU bench line (replicas + one segment on 4 GPUs).
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r2j_bench_4gpu.json 2> gpurun_out/r2j_bench_4gpu.err
cut -c1-200 gpurun_out/r2j_bench_4gpu.json; tail -n 3 gpurun_out/r2j_bench_4gpu.err
