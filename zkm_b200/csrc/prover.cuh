// prove_system: the STARK prover over a System of tables (prover.cu).
#pragma once
#include "aux.cuh"
#include <vector>
#include <functional>

namespace zkm {

struct StarkCfg {                 // StarkConfig / FriConfig (reference prover/src/config.rs:4-29)
    unsigned rate_bits = 2, cap_height = 4, pow_bits = 16, num_queries = 37, num_challenges = 2, arity_bits = 4, final_poly_bits = 5;
};
struct TableInput {
    DevBuf values;                // ncols x 2^log_n trace values on H, column-major, on the device (consumed)
    int ncols = 0, log_n = 0;
    cudaEvent_t ready = nullptr;  // if set: recorded on the copy stream once `values` is filled; the prover waits on it
    std::function<void()> wait_recorded;   // if set: blocks the host until `ready` has been recorded (uploader thread)
};
struct PublicInputs {             // PublicValues (proof.rs:52-61)
    uint32_t roots_before[8], roots_after[8];
    std::vector<uint8_t> userdata;
};

// Returns the proof in the flat u64 layout documented in include/zkm_b200.h.
std::vector<u64> prove_system(int system_id, const StarkCfg& cfg, std::vector<TableInput>& inputs, const PublicInputs& pv);

}  // namespace zkm
