#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full captures. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-run}
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_u20.json 2> gpurun_out/${TAG}_bench_u20.err
timeout 600 python bench.py --workload N22 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n22.json 2> gpurun_out/${TAG}_bench_n22.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_target.py --cols 0 --prove 20 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass' -c 10 \
    -o gpurun_out/${TAG}_ntt -f python tools/prof_target.py --cols 13 --logn 22 > gpurun_out/${TAG}_ncu_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lde_leaf_hash|merkle_level' -c 4 \
    -o gpurun_out/${TAG}_hash -f python tools/prof_target.py --cols 54 --logn 20 > gpurun_out/${TAG}_ncu_hash.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'quotient_kernel|open_segments|fri_reduce' -c 12 \
    -o gpurun_out/${TAG}_prove -f python tools/prof_target.py --cols 0 --prove 18 > gpurun_out/${TAG}_ncu_prove.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_bench_u20.json
