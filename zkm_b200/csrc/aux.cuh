// Host launchers of the STARK-layer kernels: auxiliary columns (ctl.cu), quotient (quotient.cu),
// openings (openings.cu) and FRI (fri.cu).
#pragma once
#include "devprog.cuh"
#include "batch.cuh"

namespace zkm {

constexpr int MAX_CHALLENGES = 4;
struct AuxChallenges { u64 beta[MAX_CHALLENGES]; u64 gamma[MAX_CHALLENGES]; int count; };

// aux (num_aux x n, column-major) <- lookup helper/Z columns, CTL helper columns, CTL Z columns.
void compute_aux_columns(const DProgram& prog, const tables::TableLayout& L, const u64* d_values, int log_n, const AuxChallenges& ch,
                         u64* d_aux, cudaStream_t s);

// compute_quotient_polys (prover.rs:645-789) up to (not including) the coset iNTT:
// q[a*2n + i] = (sum_k alpha_a^.. C_k(x_i)) / Z_H(x_i) on x_i = 7 w_{2n}^i, natural order.
void compute_quotient_values(int kind, const DProgram& prog, const tables::TableLayout& L, const Batch& trace, const Batch& aux,
                             const AuxChallenges& ch, const u64* alphas, int num_alphas, u64* d_q, cudaStream_t s);

// out[(c*npoints + p)*2 ..] = sum_i coeffs[c][i] * z_p^i  (StarkOpeningSet::new, proof.rs:299-334)
void eval_polys_at_points(const u64* d_coeffs, int ncols, int log_n, const gl2* points, int npoints, u64* h_out, cudaStream_t s);

}  // namespace zkm
