#!/bin/bash
# GPU-box visit for profiles only: launch list of one U20 prove, ncu --set full of the NTT passes, the hashing kernels
# and the quotient/openings kernels at production size.  Raw/details pages are exported to CSV on the box (the
# .ncu-rep files are too large to travel: gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
( cd tools/micro && ./pipe_bench ) > $O/${TAG}_pipe_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
    python tools/prof_target.py --cols 0 --prove 20 > $O/${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, count, target args...
    local name=$1 re=$2 cnt=$3; shift 3
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$re" -c $cnt -o /tmp/${TAG}_$name -f \
        python tools/prof_target.py "$@" > $O/${TAG}_ncu_$name.log 2>&1
    ncu -i /tmp/${TAG}_$name.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_$name.ncu-rep --page details --csv > $O/${TAG}_${name}_details.csv 2>/dev/null
    ls -la /tmp/${TAG}_$name.ncu-rep
}
cap ntt 'ntt_pass' 10 --cols 13 --logn 22
cap hash 'lde_leaf_hash|merkle_level' 3 --cols 54 --logn 20
cap prove 'quotient_kernel|open_segments|fri_reduce' 9 --cols 0 --prove 18
cp /tmp/${TAG}_hash.ncu-rep $O/ 2>/dev/null
du -sh $O
