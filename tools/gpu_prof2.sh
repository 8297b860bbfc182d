#!/bin/bash
# Refresh of the ncu --set full captures of the two dominant kernel families (current Poseidon v9 leaf hashing, NTT passes).
set -u
mkdir -p gpurun_out
TAG=${1:-prof2}
O=gpurun_out
cap() {
    local name=$1 re=$2 cnt=$3; shift 3
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -o /tmp/${TAG}_$name -f \
        python tools/prof_target.py "$@" > $O/${TAG}_ncu_$name.log 2>&1
    ncu -i /tmp/${TAG}_$name.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
    ls -la /tmp/${TAG}_$name.ncu-rep
}
cap hash 'lde_leaf_hash|merkle_level' 3 --cols 54 --logn 20
cap ntt 'ntt_pass' 6 --cols 13 --logn 22
du -sh $O
