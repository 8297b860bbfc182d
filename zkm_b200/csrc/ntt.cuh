// Host-side launchers for the batched Goldilocks NTT kernels (see ntt.cu).
#pragma once
#include <cuda_runtime.h>
#include "gl.cuh"

namespace zkm {

struct NttTables {
    struct Impl;
    Impl* impl;
    NttTables();
    ~NttTables();
    NttTables(const NttTables&) = delete;
    NttTables& operator=(const NttTables&) = delete;
};

// Cached two-level power table of w_{2^log_n} (or its inverse) for exponents < 2^log_n; lives as long as the tables.
struct PowTable;
PowTable ntt_root_table(NttTables& t, int log_n, int inverse, cudaStream_t s);

// All buffers are column-major: column c starts at base + c*col_stride (elements); in == out allowed.
// out[k] = sum_i in[i] w_n^(ik), natural order both sides.
void ntt_forward(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s);
// plonky2 ifft: values on H_n (natural order) -> coefficients.
void ntt_inverse(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s);
// plonky2 lde(rate_bits).coset_fft(7): n coefficients -> values on 7*H_{n<<rate_bits}, stored
// coset-major: natural LDE index m = (i << rate_bits) | j  is at  lde[c*out_cs + j*n + i].
// shift_exp_bits = e selects the coset shift 7^(2^e) (FRI round r commits on shift 7^(16^r): e = 4r).
// [coset_begin, coset_begin + coset_count) restricts the call to those cosets j (in-segment sharding, shard.cuh: a rank
// computes only the cosets it owns; the other quarters of `lde` are left untouched); coset_count < 0 = all remaining.
void lde_coset(NttTables& t, const u64* coeffs, size_t in_cs, u64* lde, size_t out_cs, int ncols, int log_n, int rate_bits,
               cudaStream_t s, int shift_exp_bits = 0, int coset_begin = 0, int coset_count = -1);
// plonky2 coset_ifft(7): values on 7*H_n (natural order) -> coefficients.
void coset_intt(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s);

}  // namespace zkm
