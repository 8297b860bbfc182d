#!/bin/bash
# Round-2 visit F (4 GPUs): NTT parity with the bulk-copy twiddle staging, in-segment sharding parity on 2 and 4 GPUs, bench at N = 4 and 2.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus.txt
( timeout 600 python -m pytest tests/test_gpu_commit.py tests/test_golden.py -m gpu -x -q -k "not full_size" 2>&1 | tail -6 ) > gpurun_out/r2f_pytest_commit.log
( timeout 600 python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "generated_on_the_device or prove_with_ops or timing_scopes" 2>&1 | tail -8 ) > gpurun_out/r2f_pytest_gen.log
( timeout 900 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -30 ) > gpurun_out/r2f_pytest_shard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2f_bench_4gpu.json 2> gpurun_out/r2f_bench_4gpu.err
CUDA_VISIBLE_DEVICES=0,1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err
tail -n 4 gpurun_out/r2f_pytest_commit.log gpurun_out/r2f_pytest_gen.log gpurun_out/r2f_pytest_shard.log
cut -c1-200 gpurun_out/r2f_bench_4gpu.json; tail -n 3 gpurun_out/r2f_bench_4gpu.err
cut -c1-200 gpurun_out/r2f_bench_2gpu.json; tail -n 3 gpurun_out/r2f_bench_2gpu.err
