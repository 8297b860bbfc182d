#!/bin/bash
# Multi-GPU visit: bench.py at N ranks (one segment per GPU, proofs gathered over NCCL).
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/multi_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $O/multi_bench_n$N.json 2> $O/multi_bench_n$N.err
tail -3 $O/multi_bench_n$N.err; python - <<PY
import json
for l in open("$O/multi_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=", d["n_gpus"], "value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
