#!/bin/bash
# Round-2 visit D (1 GPU): new tests (device-side generators, timings, prove_with_ops), corrected gl-mul variant 4, default bench.
set -u
mkdir -p gpurun_out
FAST='not benchmark_config and not full_size'
( timeout 900 python -m pytest tests -m gpu -x -q -k "$FAST" 2>&1 | tail -15 ) > gpurun_out/r2d_pytest.log
( ZKM_B200_LIB_TAG=m4 timeout 900 python -m pytest tests -m gpu -x -q -k "$FAST" 2>&1 | tail -15 ) > gpurun_out/r2d_pytest_m4.log
ZKM_B200_LIB_TAG=m4 timeout 600 python bench.py --workload N22 --steps 3 --warmup 3 > gpurun_out/r2d_bench_n22_m4.json 2> gpurun_out/r2d_bench_n22_m4.err
ZKM_B200_LIB_TAG=m4 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pageable > gpurun_out/r2d_bench_m4.json 2> gpurun_out/r2d_bench_m4.err
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_base.json 2> gpurun_out/r2d_bench_base.err
tail -n 6 gpurun_out/r2d_pytest.log gpurun_out/r2d_pytest_m4.log
cut -c1-300 gpurun_out/r2d_bench_base.json; tail -n 3 gpurun_out/r2d_bench_base.err
