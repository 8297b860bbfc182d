#!/bin/bash
# Round-2 visit W (1 GPU): the splitter (split_segment minus the step loop).
set -u
O=gpurun_out; mkdir -p $O
( timeout 300 python -m pytest tests/test_page_hash.py tests/test_abi.py -q 2>&1 | tail -12 ) > $O/r2w_pytest.log; tail -12 $O/r2w_pytest.log
