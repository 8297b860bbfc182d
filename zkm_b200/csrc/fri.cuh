// Host launchers of the FRI kernels (fri.cu).
#pragma once
#include "aux.cuh"
#include <vector>

namespace zkm {

// d_out: 6 columns of n u64 (R0.a R0.b R1.a R1.b R2.a R2.b) in coefficient form.
void fri_reduce_batches(const Batch& trace, const Batch& aux, const Batch& quot, int zstart, const std::vector<gl2>& apow, u64* d_out,
                        cudaStream_t s);
// d_r: the 6 columns evaluated on 7*H_n (natural order); d_out: 2 columns of n (F on 7*H_n).
void fri_combine(const u64* d_r, int log_n, gl2 zeta, gl2 zeta_next, gl2 v0, gl2 v1, gl2 v2, gl2 a0, gl2 a1, u64* d_out, cudaStream_t s);
// 16-value leaves of one commit-phase tree from a 2-column coset-major LDE.
void fri_leaf_rows(const u64* d_lde, size_t col_stride, int log_nr, int rate_bits, int arity_bits, u64* d_rows, cudaStream_t s);
// d_in: 2 columns of n_in coefficients -> d_out: 2 columns of n_in >> arity_bits.
void fri_fold(const u64* d_in, size_t n_in, int arity_bits, gl2 beta, u64* d_out, cudaStream_t s);
// Minimum witness w with >= min_lz leading zeros in state'[7] after placing w at `pos` and permuting.
u64 fri_pow_grind(const u64 state[12], int pos, int min_lz, cudaStream_t s);

}  // namespace zkm
