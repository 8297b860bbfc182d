#!/bin/bash
# Round-2 visit A: parity tests incl. the U20 / 2^22 oracle comparisons, both bench arms, quotient register-budget A/B.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > gpurun_out/r2a_pytest.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_u20.json 2> gpurun_out/r2a_bench_u20.err
ZKM_B200_LIB_TAG=q2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_u20_q2.json 2> gpurun_out/r2a_bench_u20_q2.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --host-memory pageable > gpurun_out/r2a_bench_u20_pageable.json 2> gpurun_out/r2a_bench_u20_pageable.err
tail -5 gpurun_out/r2a_pytest.log
cut -c1-600 gpurun_out/r2a_bench_u20.json
