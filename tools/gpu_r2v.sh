#!/bin/bash
# Round-2 visit V (1 GPU): the fast part of the GPU suite on the final library (everything but the full-size / benchmark-config tests).
set -u
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -x -q -k "not benchmark_config and not full_size and not sha2_guest and not 2_21_rows" 2>&1 | tail -6 ) > $O/r2v_pytest_fast.log 2>&1
tail -8 $O/r2v_pytest_fast.log
