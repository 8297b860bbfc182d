#!/bin/bash
# Round-2 visit G (1 GPU, short): parity of the commit slice and the fast suite after the NTT shared-memory opt-in fix.
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_commit.py tests/test_golden.py -m gpu -x -q -k "not full_size" 2>&1 | tail -6 ) > gpurun_out/r2g_pytest_commit.log
( timeout 600 python -m pytest tests -m gpu -x -q -k "not benchmark_config and not full_size and not commit_matches and not ntt_matches" 2>&1 | tail -12 ) > gpurun_out/r2g_pytest_fast.log
timeout 300 python bench.py --workload N22 --steps 3 --warmup 3 > gpurun_out/r2g_bench_n22.json 2> gpurun_out/r2g_bench_n22.err
tail -n 5 gpurun_out/r2g_pytest_commit.log gpurun_out/r2g_pytest_fast.log; tail -n 1 gpurun_out/r2g_bench_n22.json | cut -c1-250
