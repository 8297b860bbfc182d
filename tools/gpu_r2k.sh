#!/bin/bash
# Round-2 visit K: the new device-side generators (six hash tables), page hashing, rate_bits = 3 commitments.
set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_tracegen.py tests/test_page_hash.py "tests/test_gpu_commit.py::test_commit_at_the_recursion_blow_up" \
    "tests/test_gpu_prove.py::test_logic_and_poseidon_tables_generated_on_the_device" -m gpu -q 2>&1 | tail -60 ) > $O/r2k_pytest.log
tail -60 $O/r2k_pytest.log
