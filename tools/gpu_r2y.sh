#!/bin/bash
# Round-2 visit Y (1 GPU): Poseidon S-box with the PTX product on some of its four multiplies (ZKM_P9_MULMIX), micro-benchmark only.
set -u
O=gpurun_out; mkdir -p $O
for m in 0 2 8 10 5 15 0; do ( cd tools/micro && ./poseidon_bench_mix$m quick ); done > $O/r2y_poseidon_mulmix.txt 2>&1
cat $O/r2y_poseidon_mulmix.txt
