#!/bin/bash
# Round-2 visit S (1 GPU): final bench line, launch list of one bench step, ncu --set full of the NTT passes and the hash kernels.
# Only small text files stay under gpurun_out/ (the .ncu-rep captures are summarised on the box and deleted: 64 MiB cap).
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2s_bench_u20.json 2> $O/r2s_bench_u20.err
timeout 300 python bench.py --workload N22 --steps 3 --warmup 3 > $O/r2s_bench_n22.json 2> $O/r2s_bench_n22.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2s_launches.csv \
    python bench.py --steps 1 --warmup 1 --workers 1 --no-cpu-baseline --no-pageable > $O/r2s_launches.log 2>&1
mkdir -p /tmp/ncu
cap() {  # name, kernel regex, count, target args...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/ncu/$name -f python tools/prof_target.py "$@" > $O/r2s_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep $O/r2s_${name}_ncu_full.csv >> $O/r2s_ncu_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > /tmp/ncu/$name.raw.csv 2>/dev/null
  python tools/ncu_pick.py /tmp/ncu/$name.raw.csv > $O/r2s_${name}_ncu_pick.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
cap ntt 'ntt_pass' 10 --cols 13 --logn 22
cap hash 'lde_leaf_hash|merkle_level_kernel' 4 --cols 54 --logn 20
cut -c1-400 $O/r2s_bench_u20.json; echo; tail -2 $O/r2s_bench_u20.err; head -5 $O/r2s_ntt_ncu_pick.txt | cut -c1-200; wc -l $O/r2s_launches.csv
