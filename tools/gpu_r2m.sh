#!/bin/bash
# Round-2 visit M: pageable-source upload through the pinned bounce ring, A/B against the driver's own pageable path.
set -u
O=gpurun_out; mkdir -p $O
ZKM_BOUNCE=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r2m_bench_bounce0.json 2> $O/r2m_bench_bounce0.err
ZKM_BOUNCE=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r2m_bench_bounce1.json 2> $O/r2m_bench_bounce1.err
for f in $O/r2m_bench_bounce0.json $O/r2m_bench_bounce1.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'pageable', d['e2e'].get('pageable_host_memory'), 'single', d.get('single_proof_latency_ms'))
PY
done
tail -3 $O/r2m_bench_bounce1.err
( timeout 600 python -m pytest tests/test_gpu_prove.py -m gpu -q -k "row_major or drop_in or grouped or valid_trace_proof" 2>&1 | tail -5 ) > $O/r2m_pytest.log; tail -5 $O/r2m_pytest.log
