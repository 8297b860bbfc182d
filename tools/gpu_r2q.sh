#!/bin/bash
# Round-2 visit Q (2 GPUs): blocking host waits (ZKM_BLOCKING_SYNC) A/B -- 1-GPU and 2-GPU bench lines on the same box.
set -u
O=gpurun_out; mkdir -p $O
echo "nproc $(nproc) affinity $(python -c 'import os; print(len(os.sched_getaffinity(0)))') load $(cat /proc/loadavg)" > $O/r2q_host.txt
for v in 0 1; do
  ZKM_BLOCKING_SYNC=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-pageable > $O/r2q_bench_1gpu_block$v.json 2> $O/r2q_bench_1gpu_block$v.err
  ZKM_BLOCKING_SYNC=$v timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$v bench.py --gpus 2 --steps 4 --warmup 3 --no-in-segment > $O/r2q_bench_2gpu_block$v.json 2> $O/r2q_bench_2gpu_block$v.err
done
cat $O/r2q_host.txt
python - <<'PY'
import json
for n in (1,2):
    for v in (0,1):
        try:
            d=json.loads(open(f'gpurun_out/r2q_bench_{n}gpu_block{v}.json').read().strip().splitlines()[-1])
            print(n,'GPU blocking',v,'value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'single',round(d['single_proof_latency_ms'],1))
        except Exception as e:
            print(n,v,'failed',e)
PY
