#!/bin/bash
# Round-2 visit X (1 GPU): 3 vs 4 proofs in flight per GPU on the final library.
set -u
O=gpurun_out; mkdir -p $O
for w in 3 4; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable --workers $w > $O/r2x_bench_w$w.json 2> $O/r2x_bench_w$w.err
done
python - <<'PY'
import json
for w in (3,4):
    d=json.loads(open(f'gpurun_out/r2x_bench_w{w}.json').read().strip().splitlines()[-1])
    print('workers',w,'value',round(d['value'],3),'e2e',round(d['e2e']['value'],3))
PY
