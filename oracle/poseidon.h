// ORACLE (test infrastructure): Goldilocks Poseidon permutation (width 12, rate 8, 4+22+4 rounds,
// x^7) and the plonky2 hashing modes built on it.
//  * permutation, naive schedule: restates the round structure of reference
//    prover/src/poseidon/poseidon_stark.rs:51-95 (full/partial/full), constant layer :158-163,
//    S-box x^7 :210-243, MDS row (circulant + diagonal) :310-343, with parameters from
//    prover/src/poseidon/constants.rs:11-105 (oracle/poseidon_consts.h is generated from it).
//  * permutation, fast partial-round schedule: poseidon_stark.rs:76-95,389-400,463-501 with
//    constants.rs:107-870.  tests check naive == fast == the known answers of SURVEY Appendix D.
//  * hash_no_pad / hash_or_noop / two_to_one: plonky2 0.1.4 hash/hashing.rs, hash/poseidon.rs
//    (not in /root/reference; restated from SURVEY Appendix A.4).
#pragma once
#include "field.h"
#include "poseidon_consts.h"
#include <array>
#include <cstring>

namespace orc {

typedef std::array<Fp, 12> PState;
struct Digest {
    Fp e[4];
    bool operator==(const Digest& o) const {
        return e[0] == o.e[0] && e[1] == o.e[1] && e[2] == o.e[2] && e[3] == o.e[3];
    }
};

static inline Fp sbox7(Fp x) {
    Fp x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    return x3 * x4;
}

// res[r] = sum_i state[(i+r)%12]*CIRC[i] + state[r]*DIAG[r]     (poseidon_stark.rs:331-343)
static inline void mds_layer(PState& s) {
    PState out;
    for (int r = 0; r < 12; r++) {
        u128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12].v * POSEIDON_MDS_CIRC[i];
        acc += (u128)s[r].v * POSEIDON_MDS_DIAG[r];
        out[r].v = gl_reduce128(acc);
    }
    s = out;
}

static inline void poseidon_naive(PState& s) {
    int round = 0;
    for (int r = 0; r < 4; r++, round++) {
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
    }
    for (int r = 0; r < 22; r++, round++) {
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
        s[0] = sbox7(s[0]);
        mds_layer(s);
    }
    for (int r = 0; r < 4; r++, round++) {
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
    }
}

// Fast partial rounds (poseidon_stark.rs:76-95): partial_first_constant_layer,
// mds_partial_layer_init, then 22 x { sbox(s0); s0 += c_i; mds_partial_layer_fast }.
static inline void poseidon_fast(PState& s) {
    int round = 0;
    for (int r = 0; r < 4; r++, round++) {
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
    }
    for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
    {   // mds_partial_layer_init: result[0] = s[0]; result[c+1] = sum_r s[r+1]*M[r][c]
        PState out;
        out[0] = s[0];
        for (int c = 0; c < 11; c++) {
            Fp acc;
            for (int r = 0; r < 11; r++)
                acc += s[r + 1] * Fp(POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r * 11 + c]);
            out[c + 1] = acc;
        }
        s = out;
    }
    for (int i = 0; i < 22; i++) {
        s[0] = sbox7(s[0]);
        s[0] += Fp(POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[i]);
        // mds_partial_layer_fast: d = s0*(CIRC[0]+DIAG[0]) + sum_{j>=1} s_j*W_HAT[i][j-1];
        //                         s_j += s0*VS[i][j-1]
        Fp d = s[0] * Fp(POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0]);
        for (int j = 1; j < 12; j++) d += s[j] * Fp(POSEIDON_FAST_PARTIAL_ROUND_W_HATS[i * 11 + j - 1]);
        PState out;
        out[0] = d;
        for (int j = 1; j < 12; j++) out[j] = s[j] + s[0] * Fp(POSEIDON_FAST_PARTIAL_ROUND_VS[i * 11 + j - 1]);
        s = out;
    }
    round += 22;
    for (int r = 0; r < 4; r++, round++) {
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
    }
}

// Optimised CPU schedule used for everything but the cross-checks (tests assert naive == fast == opt ==
// Appendix D): the same permutation with (i) the MDS row sums taken over 32-bit halves in u64 accumulators
// (the circulant entries are < 64, so no 128-bit arithmetic and the loops auto-vectorise), one reduction per
// output, and (ii) lazily reduced S-box products.  This is the CPU arm's hash; a scalar Rust/plonky2 build
// uses the same two ideas (plonky2 hash/poseidon_goldilocks.rs).
static inline void mds_layer_opt(u64* s) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 lo[24], hi[24];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = s[i] & 0xFFFFFFFFULL; hi[i] = hi[i + 12] = s[i] >> 32; }
    u64 al[12], ah[12];
    for (int r = 0; r < 12; r++) {
        u64 a = 0, b = 0;
        for (int i = 0; i < 12; i++) { a += lo[i + r] * C[i]; b += hi[i + r] * C[i]; }
        al[r] = a; ah[r] = b;
    }
    al[0] += lo[0] * 8; ah[0] += hi[0] * 8;
    for (int r = 0; r < 12; r++) {
        u128 v = (u128)al[r] + ((u128)ah[r] << 32);
        s[r] = gl_reduce128(v);
    }
}
static inline u64 mulred(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static inline u64 sbox7_opt(u64 x) {
    u64 x2 = mulred(x, x), x3 = mulred(x2, x), x4 = mulred(x2, x2);
    return mulred(x3, x4);
}
static inline void poseidon_opt(PState& st) {
    u64 s[12];
    for (int i = 0; i < 12; i++) s[i] = st[i].v;
    int rc = 0;
    for (int r = 0; r < 30; r++) {
        const bool full = r < 4 || r >= 26;
        for (int i = 0; i < 12; i++) {
            u64 t = s[i] + POSEIDON_ALL_ROUND_CONSTANTS[rc + i];          // both < p: at most one wrap
            if (t < s[i] || t >= GL_P) t -= GL_P;
            s[i] = t;
        }
        rc += 12;
        if (full) { for (int i = 0; i < 12; i++) s[i] = sbox7_opt(s[i]); }
        else s[0] = sbox7_opt(s[0]);
        mds_layer_opt(s);
    }
    for (int i = 0; i < 12; i++) st[i].v = s[i];
}

static inline void poseidon(PState& s) { poseidon_opt(s); }

// plonky2 hash_n_to_m_no_pad with m = 4: overwrite-mode sponge, rate 8 (Appendix A.4).
static inline Digest hash_no_pad(const Fp* in, size_t n) {
    PState s;
    for (size_t off = 0; off < n; off += 8) {
        size_t k = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < k; i++) s[i] = in[off + i];
        poseidon(s);
    }
    Digest d;
    for (int i = 0; i < 4; i++) d.e[i] = s[i];
    return d;
}
// plonky2 Hasher::hash_or_noop: <= 4 elements are copied (zero padded), else hash_no_pad.
static inline Digest hash_or_noop(const Fp* in, size_t n) {
    if (n <= 4) {
        Digest d;
        for (size_t i = 0; i < n; i++) d.e[i] = in[i];
        return d;
    }
    return hash_no_pad(in, n);
}
// plonky2 compress / two_to_one.
static inline Digest two_to_one(const Digest& l, const Digest& r) {
    PState s;
    for (int i = 0; i < 4; i++) { s[i] = l.e[i]; s[4 + i] = r.e[i]; }
    poseidon(s);
    Digest d;
    for (int i = 0; i < 4; i++) d.e[i] = s[i];
    return d;
}

}  // namespace orc
