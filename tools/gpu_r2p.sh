#!/bin/bash
# Round-2 visit P (2 GPUs): in-segment sharding parity after the NTT full-table change; 2-GPU bench line.
set -u
O=gpurun_out; mkdir -p $O
( timeout 500 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -12 ) > $O/r2p_pytest_shard.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 4 --warmup 3 > $O/r2p_bench_2gpu.json 2> $O/r2p_bench_2gpu.err
tail -n 5 $O/r2p_pytest_shard.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'single',d['single_proof_latency_ms'],'in_segment',d.get('in_segment'))
PY
tail -n 3 $O/r2p_bench_2gpu.err
