// prove_system: the STARK prover over a System of tables (prover.cu).
#pragma once
#include "aux.cuh"
#include <vector>
#include <functional>
#include <algorithm>
#include <numeric>

namespace zkm {

struct StarkCfg {                 // StarkConfig / FriConfig (reference prover/src/config.rs:4-29)
    unsigned rate_bits = 2, cap_height = 4, pow_bits = 16, num_queries = 37, num_challenges = 2, arity_bits = 4, final_poly_bits = 5;
};
struct TableInput {
    DevBuf values;                // ncols x 2^log_n trace values on H, column-major, on the device (consumed)
    int ncols = 0, log_n = 0;
    cudaEvent_t ready = nullptr;  // if set: recorded on the copy stream once `values` is filled; the prover waits on it
    std::function<void()> wait_recorded;   // if set: blocks the host until `ready` has been recorded (uploader thread)
    // large host tables arrive in column groups: group k = columns [group_ends[k-1], group_ends[k]) is complete once
    // group_ready[k] (recorded on the copy stream) has fired; wait_group(k) blocks the host until that event is recorded
    std::vector<int> group_ends;
    std::vector<cudaEvent_t> group_ready;
    std::function<void(size_t)> wait_group;
};
struct PublicInputs {             // PublicValues (proof.rs:52-61)
    uint32_t roots_before[8], roots_after[8];
    std::vector<uint8_t> userdata;
};

// Order in which the tables are uploaded and committed: ascending size, ties in table order.
inline std::vector<size_t> commit_order(const std::vector<size_t>& sizes) {
    std::vector<size_t> o(sizes.size());
    std::iota(o.begin(), o.end(), (size_t)0);
    std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return sizes[a] < sizes[b]; });
    return o;
}

// Returns the proof in the flat u64 layout documented in include/zkm_b200.h.
std::vector<u64> prove_system(int system_id, const StarkCfg& cfg, std::vector<TableInput>& inputs, const PublicInputs& pv);

// One table's stage outputs under given challenges (prover.cu): auxiliary columns, quotient coefficients, openings.
void stage_single_table(int system_id, int table_index, const StarkCfg& cfg, DevBuf&& values, int ncols, int log_n, const AuxChallenges& ctl_ch,
                        const u64* alphas, gl2 zeta, std::vector<u64>& aux_out, std::vector<u64>& quot_out, std::vector<u64>& open_out);

}  // namespace zkm
