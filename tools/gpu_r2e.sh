#!/bin/bash
# Round-2 visit E (1 GPU): the complete GPU suite on the product build, then both bench arms as the driver runs them.
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -16 ) > gpurun_out/r2e_pytest.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err ) 2> gpurun_out/r2e_time_ref.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_u20.json 2> gpurun_out/r2e_bench_u20.err ) 2> gpurun_out/r2e_time_u20.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
tail -n 12 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_time_ref.txt gpurun_out/r2e_time_u20.txt; tail -n 2 gpurun_out/r2e_smoke.log
cut -c1-300 gpurun_out/r2e_bench_u20.json
