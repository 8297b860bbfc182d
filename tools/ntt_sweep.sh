#!/bin/bash
# NTT tile/CTA-size sweep on the GPU box: N20 commit (54 columns x 2^20), NTT family time per setting.
O=gpurun_out; mkdir -p $O
for cfg in "0 0" "8 512" "8 256" "4 1024" "4 512" "4 256" "2 512" "2 256" "16 1024"; do
  set -- $cfg
  ZKM_NTT_T=$1 ZKM_NTT_THREADS=$2 timeout 120 python bench.py --workload N20 --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if d['config']['columns']==54: print('T=$1 threads=$2', 'ntt ms', round(d['kernel_families']['ntt_pass']['ms_per_step'],3), 'launches', d['kernel_families']['ntt_pass']['launches_per_step'])
    elif 'rror' in l: print(l.strip()[:200])
"
done | tee $O/ntt_sweep.txt
