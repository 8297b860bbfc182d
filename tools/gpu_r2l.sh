#!/bin/bash
# Round-2 visit L: stage-level parity (aux columns, quotient coefficients, openings) against the oracle.
set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_stages.py -m gpu -q 2>&1 | tail -40 ) > $O/r2l_pytest.log
tail -40 $O/r2l_pytest.log
