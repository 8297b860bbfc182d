#!/bin/bash
# Round-2 visit T (1 GPU): L2 prefetch of the factor-table sectors in NTT pass A -- N22 sweep, U20 bench, commit parity.
set -u
O=gpurun_out; mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_commit.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3 ) > $O/r2t_pytest_commit.log; tail -2 $O/r2t_pytest_commit.log
timeout 300 python bench.py --workload N22 --steps 3 --warmup 3 > $O/r2t_bench_n22.json 2> $O/r2t_bench_n22.err
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable > $O/r2t_bench_u20.json 2> $O/r2t_bench_u20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench_n22.json').read().strip().splitlines()[-1])
print('N22', round(d['value'],1), {k:(round(x['ms_per_step'],2)) for k,x in d['kernel_families'].items()})
d=json.loads(open('gpurun_out/r2t_bench_u20.json').read().strip().splitlines()[-1])
print('U20 value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'single',round(d['single_proof_latency_ms'],1),{k:round(x['ms_per_step'],2) for k,x in d['kernel_families'].items() if k in ('ntt_pass','leaf_hash')}, d.get('roofline_ntt',{}).get('achieved'))
PY
