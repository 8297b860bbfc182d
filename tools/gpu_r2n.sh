#!/bin/bash
# Round-2 visit N: NTT pass A with full-size factor tables (ZKM_NTT_FULLTAB), parity + A/B on the N22 sweep and the U20 proof.
set -u
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_commit.py tests/test_golden.py tests/test_stages.py -m gpu -x -q 2>&1 | tail -6 ) > $O/r2n_pytest_commit.log
tail -4 $O/r2n_pytest_commit.log
for v in 0 1; do
  ZKM_NTT_FULLTAB=$v timeout 300 python bench.py --workload N22 --steps 3 --warmup 3 > $O/r2n_bench_n22_fulltab$v.json 2> $O/r2n_bench_n22_fulltab$v.err
  ZKM_NTT_FULLTAB=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable > $O/r2n_bench_u20_fulltab$v.json 2> $O/r2n_bench_u20_fulltab$v.err
done
python - <<'PY'
import json
for v in (0,1):
    d=json.loads(open(f'gpurun_out/r2n_bench_n22_fulltab{v}.json').read().strip().splitlines()[-1])
    print('N22 fulltab',v, json.dumps(d)[:600])
    d=json.loads(open(f'gpurun_out/r2n_bench_u20_fulltab{v}.json').read().strip().splitlines()[-1])
    print('U20 fulltab',v,'value',round(d['value'],3),'single',round(d['single_proof_latency_ms'],1),{k:round(x['ms_per_step'],2) for k,x in d['kernel_families'].items() if k in ('ntt_pass','leaf_hash','merkle_levels','quotient')}, d.get('roofline_ntt',{}).get('achieved'))
PY
