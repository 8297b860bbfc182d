#!/bin/bash
set -u
TAG=${1:-v4}
O=gpurun_out; mkdir -p $O
timeout 300 python tools/prof_target.py --cols 0 --prove 20 --warm 2 --trace > $O/${TAG}_trace.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'k_perm_vx<\(int\)23, \(int\)8>' -c 1 -o /tmp/${TAG}_perm -f \
    tools/micro/poseidon_bench > $O/${TAG}_ncu_perm.log 2>&1
ncu -i /tmp/${TAG}_perm.ncu-rep --page raw --csv > $O/${TAG}_perm_raw.csv 2>/dev/null
grep -c zkm_b200 $O/${TAG}_trace.log; wc -c $O/${TAG}_perm_raw.csv
