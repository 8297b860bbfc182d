#!/bin/bash
# Round-2 visit U (1 GPU): the C++ host mirror proving through include/zkm_b200.hpp; smoke().
set -u
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_cpp_host.py -m gpu -q 2>&1 | tail -12 ) > $O/r2u_pytest_cpp.log; tail -12 $O/r2u_pytest_cpp.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/r2u_smoke.log; tail -3 $O/r2u_smoke.log
