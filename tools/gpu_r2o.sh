#!/bin/bash
# Round-2 visit O (1 GPU): the whole GPU suite (timed), sanitizer over the kernels added in this session.
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/r2o_pytest.log 2>&1
tail -8 $O/r2o_pytest.log
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "sha_extend_tables or sha_compress_tables or byte_sponge or page_hashing or (recursion_blow_up and 3-5) or (stage_outputs and 1) or (ntt_matches_oracle)" 2>&1 | tail -25 ) > $O/r2o_sanitizer_memcheck.log
( timeout 400 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_commit.py -m gpu -x -q -k "ntt_matches_oracle or (commit_matches_oracle and 13-10)" 2>&1 | tail -25 ) > $O/r2o_sanitizer_racecheck.log
tail -n 5 $O/r2o_sanitizer_memcheck.log $O/r2o_sanitizer_racecheck.log
