#!/bin/bash
# Round-2 visit R (2 GPUs): large-table uploads of concurrent proofs take turns on the PCIe link (ZKM_UPLOAD_FIFO) -- A/B.
set -u
O=gpurun_out; mkdir -p $O
for v in 0 1; do
  ZKM_UPLOAD_FIFO=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r2r_bench_1gpu_fifo$v.json 2> $O/r2r_bench_1gpu_fifo$v.err
  ZKM_UPLOAD_FIFO=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$v bench.py --gpus 2 --steps 4 --warmup 3 --no-in-segment > $O/r2r_bench_2gpu_fifo$v.json 2> $O/r2r_bench_2gpu_fifo$v.err
done
python - <<'PY'
import json
for n in (1,2):
    for v in (0,1):
        try:
            d=json.loads(open(f'gpurun_out/r2r_bench_{n}gpu_fifo{v}.json').read().strip().splitlines()[-1])
            print(n,'GPU fifo',v,'value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'pageable',(d['e2e'].get('pageable_host_memory') or {}).get('value'),'single',round(d['single_proof_latency_ms'],1))
        except Exception as e:
            print(n,v,'failed',e)
PY
( timeout 300 python -m pytest tests/test_gpu_prove.py -m gpu -q -k "two_worker or mixed_shape or grouped or row_major_tables" 2>&1 | tail -3 ) > $O/r2r_pytest.log; tail -3 $O/r2r_pytest.log
