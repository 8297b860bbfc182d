// Device-side table generators from operation logs: the tables whose rows are a pure function of a
// compact log of operations are built directly in HBM, column-major, so that only the log crosses PCIe.
//   Logic     reference logic.rs:108-183  (Operation::into_row, generate_trace_rows: one row per operation, zero padding)
//             log entry = 3 words: operator (0 AND, 1 OR, 2 XOR, 3 NOR), input0, input1 (u32 each) -> 69 columns (23x smaller)
//   Poseidon  reference poseidon/poseidon_stark.rs:51-95 (poseidon_with_witness), :105-145 (generate_trace_rows[_for_perm]:
//             one permutation per row with the x^3 / x^7 witness of every S-box; padding rows = the permutation of zero with
//             FILTER = 0)  log entry = 13 words: the 12 input elements (canonical), timestamp -> 262 columns (20x smaller)
// The Memory table's generator lives in memtrace.cu.  One thread per row; every store is coalesced across the warp.
#include "dev.cuh"
#include "tables/logic.h"
#include "tables/poseidon.h"

namespace zkm {

static size_t padded_rows(size_t n_ops, size_t min_rows) {
    size_t n = n_ops > min_rows ? n_ops : min_rows, p = 1;
    while (p < n) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------------------ Logic
__global__ void logic_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace lg = tables::logic;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 op = 4, a = 0, b = 0;
    if (r < n_ops) {
        op = ops[3 * r]; a = ops[3 * r + 1]; b = ops[3 * r + 2];
        if (op > 3 || a >> 32 || b >> 32) { atomicExch(bad, 1u); op = 4; a = b = 0; }
    }
    const u32 x = (u32)a, y = (u32)b;
    const u32 res = op == 0 ? (x & y) : op == 1 ? (x | y) : op == 2 ? (x ^ y) : op == 3 ? ~(x | y) : 0u;
    cols[(size_t)lg::IS_AND * n + r] = op == 0; cols[(size_t)lg::IS_OR * n + r] = op == 1;
    cols[(size_t)lg::IS_XOR * n + r] = op == 2; cols[(size_t)lg::IS_NOR * n + r] = op == 3;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        cols[(size_t)(lg::INPUT0 + i) * n + r] = (x >> i) & 1;
        cols[(size_t)(lg::INPUT1 + i) * n + r] = (y >> i) & 1;
    }
    cols[(size_t)lg::RESULT * n + r] = res;
}

size_t logic_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    namespace lg = tables::logic;
    static_assert(lg::NUM_COLUMNS == 69, "logic layout");
    const size_t n = padded_rows(n_ops, min_rows);
    DevBuf ops(3 * n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, 3 * n_ops);
    flag.zero();
    cols.alloc((size_t)lg::NUM_COLUMNS * n, s);
    ProfScope ps("logic_trace", s, 24.0 * (double)n_ops + 8.0 * lg::NUM_COLUMNS * (double)n);
    logic_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ops.p, n_ops, n, cols.p, (unsigned*)flag.p);
    ZKM_LAUNCHED();
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "logic operation out of range (operator 0..3, 32-bit inputs)");
    return n;
}

// --------------------------------------------------------------------------------------------------------- Poseidon
__global__ void __launch_bounds__(128) poseidon_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols,
                                                            unsigned* bad) {
    namespace pz = tables::poseidon;
    using namespace tables::poseidon;                  // ZKM_K(name) pastes the bare table name
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    auto put = [&](int c, gl v) { cols[(size_t)c * n + row] = v.v; };
    gl st[12];
    const bool real = row < n_ops;
    u64 ts = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 v = real ? ops[13 * row + i] : 0;
        if (v >= GL_P) { atomicExch(bad, 1u); v = 0; }
        st[i] = gl(v);
        put(pz::reg_in(i), st[i]);
    }
    if (real) ts = ops[13 * row + 12];
    put(pz::FILTER, gl(real ? 1 : 0));
    put(pz::TIMESTAMP, gl(ts));
    int round = 0;
    auto full = [&](int r, bool second) {
#pragma unroll 1
        for (int i = 0; i < 12; i++) {
            gl x = st[i] + gl(ZKM_K(PT_RC)[12 * round + i]);
            gl x3 = x * x * x, x7 = x3 * x3 * x;
            const int base = second ? pz::reg_full1_s0(r, i) : pz::reg_full0_s0(r, i);
            put(base, x3); put(base + 1, x7);
            st[i] = x7;
        }
        pz::mds_layer<gl>(st);
        round++;
    };
#pragma unroll 1
    for (int r = 0; r < 4; r++) full(r, false);
    // partial rounds in the reference's fast formulation (poseidon_stark.rs:76-95,389-400,463-501)
#pragma unroll 1
    for (int i = 0; i < 12; i++) st[i] = st[i] + gl(ZKM_K(PT_FIRST)[i]);
    {
        gl o[12];
        o[0] = st[0];
#pragma unroll 1
        for (int c = 0; c < 11; c++) {
            gl acc = gl::zero();
            for (int r = 0; r < 11; r++) acc = acc + st[r + 1] * gl(ZKM_K(PT_INIT)[r * 11 + c]);
            o[c + 1] = acc;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = o[i];
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        gl x3 = st[0] * st[0] * st[0], x7 = x3 * x3 * st[0];
        put(pz::reg_partial_s0(r), x3); put(pz::reg_partial_s0(r) + 1, x7);
        st[0] = x7;
        if (r < 21) st[0] = st[0] + gl(ZKM_K(PT_PRC)[r]);
        gl d = st[0] * gl(ZKM_K(PT_CIRC)[0] + ZKM_K(PT_DIAG)[0]);
        for (int j = 1; j < 12; j++) d = d + st[j] * gl(ZKM_K(PT_WHAT)[r * 11 + j - 1]);
        for (int j = 1; j < 12; j++) st[j] = st[j] + st[0] * gl(ZKM_K(PT_VS)[r * 11 + j - 1]);
        st[0] = d;
    }
    round += 22;
#pragma unroll 1
    for (int r = 0; r < 4; r++) full(r, true);
#pragma unroll
    for (int i = 0; i < 12; i++) put(pz::reg_out(i), st[i]);
}

size_t poseidon_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    namespace pz = tables::poseidon;
    static_assert(pz::NUM_COLUMNS == 262, "poseidon layout");
    const size_t n = padded_rows(n_ops, min_rows);
    DevBuf ops(13 * n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, 13 * n_ops);
    flag.zero();
    cols.alloc((size_t)pz::NUM_COLUMNS * n, s);
    ProfScope ps("poseidon_trace", s, 104.0 * (double)n_ops + 8.0 * pz::NUM_COLUMNS * (double)n);
    poseidon_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ops.p, n_ops, n, cols.p, (unsigned*)flag.p);
    ZKM_LAUNCHED();
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "poseidon input is not a canonical field element");
    return n;
}

}  // namespace zkm
