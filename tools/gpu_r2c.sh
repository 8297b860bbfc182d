#!/bin/bash
# Round-2 visit C (1 GPU): fast parity subset, A/B builds (gl mul 4, quotient occupancy), ncu summaries, sanitizer.
# Only small text files stay under gpurun_out/ (the .ncu-rep captures are summarised on the box and deleted: 64 MiB cap).
set -u
mkdir -p gpurun_out
FAST='not benchmark_config and not full_size'
( timeout 900 python -m pytest tests -m gpu -x -q -k "$FAST" 2>&1 | tail -8 ) > gpurun_out/r2c_pytest.log
( ZKM_B200_LIB_TAG=m4 timeout 900 python -m pytest tests -m gpu -x -q -k "$FAST" 2>&1 | tail -8 ) > gpurun_out/r2c_pytest_m4.log
for tag in "" m4 q3 q4; do
  ZKM_B200_LIB_TAG=$tag timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_${tag:-base}.json 2> gpurun_out/r2c_bench_${tag:-base}.err
done
timeout 600 python bench.py --workload N22 --steps 3 --warmup 3 > gpurun_out/r2c_bench_n22.json 2> gpurun_out/r2c_bench_n22.err
ZKM_B200_LIB_TAG=m4 timeout 600 python bench.py --workload N22 --steps 3 --warmup 3 > gpurun_out/r2c_bench_n22_m4.json 2> gpurun_out/r2c_bench_n22_m4.err
# ncu: launch list of a bench run + full captures of the dominant kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches.csv \
    python bench.py --steps 1 --warmup 1 --workers 1 --no-cpu-baseline > gpurun_out/r2c_launches.log 2>&1
mkdir -p /tmp/ncu
cap() {  # name, kernel regex, count, target args...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/ncu/$name -f python tools/prof_target.py "$@" > gpurun_out/r2c_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep gpurun_out/r2c_${name}_ncu_full.csv >> gpurun_out/r2c_ncu_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > /tmp/ncu/$name.raw.csv 2>/dev/null
  python tools/ncu_pick.py /tmp/ncu/$name.raw.csv > gpurun_out/r2c_${name}_ncu_pick.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
cap hash 'lde_leaf_hash|merkle_level_kernel' 4 --cols 54 --logn 20
cap ntt 'ntt_pass' 10 --cols 13 --logn 22
cap quot 'quotient_kernel' 6 --cols 0 --prove 18
# sanitizer: memcheck + racecheck over small commit / prove tests
( timeout 500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "(commit_matches_oracle and (13-10 or 5-6 or 4-6)) or (equals_oracle_proof_and_verifies and (1 or 3 or 4)) or memory_table" 2>&1 | tail -25 ) > gpurun_out/r2c_sanitizer_memcheck.log
( timeout 500 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_commit.py -m gpu -x -q -k "commit_matches_oracle and (13-10 or 5-6)" 2>&1 | tail -25 ) > gpurun_out/r2c_sanitizer_racecheck.log
du -sh gpurun_out
tail -n 3 gpurun_out/r2c_pytest.log gpurun_out/r2c_pytest_m4.log
tail -n 4 gpurun_out/r2c_sanitizer_memcheck.log gpurun_out/r2c_sanitizer_racecheck.log
