#!/bin/bash
# Short GPU-box visit: micro-benchmarks, parity tests, one bench line.  usage: gpu_quick.sh TAG [bench args]
set -u
TAG=${1:-q}; shift
O=gpurun_out; mkdir -p $O
( cd tools/micro && ./poseidon_bench ) > $O/${TAG}_poseidon_bench.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 "$@" > $O/${TAG}_bench_u20.json 2> $O/${TAG}_bench_u20.err
cat $O/${TAG}_poseidon_bench.txt; tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench_u20.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step']); print({k:round(v['ms_per_step'],2) for k,v in d['kernel_families'].items()})"
