#!/bin/bash
set -u
TAG=${1:-v3}
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench_u20.json 2> $O/${TAG}_bench_u20.err
ZKM_TRACE=1 timeout 300 python tools/prof_target.py --cols 0 --prove 20 > $O/${TAG}_trace.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'k_perm_vx<9, ?8>' -c 1 -o /tmp/${TAG}_perm -f \
    tools/micro/poseidon_bench > $O/${TAG}_ncu_perm.log 2>&1
ncu -i /tmp/${TAG}_perm.ncu-rep --page raw --csv > $O/${TAG}_perm_raw.csv 2>/dev/null
tail -3 $O/${TAG}_pytest.log; python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_u20.json").read()); print(d['ms_per_step'], d['e2e']['ms_per_step']); print({k:round(v['ms_per_step'],2) for k,v in d['kernel_families'].items()})
PY
