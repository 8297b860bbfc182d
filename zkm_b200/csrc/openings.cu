// Polynomial openings: every committed polynomial evaluated at a few extension-field points from its
// coefficients.  Replaces StarkOpeningSet::new's per-polynomial Horner loops (reference
// prover/src/proof.rs:299-334: eval at zeta and g*zeta for trace/aux, at zeta for the quotient chunks,
// at 1 for the CTL Z polynomials).
// One pass over the coefficients for all points: a CTA owns a 2^SEG_BITS-coefficient segment of one
// column, each thread runs Horner over 2^THREAD_BITS consecutive coefficients, the CTA combines the
// partial values with a power tree (z^(2^k) precomputed on the host), and a second tiny kernel adds the
// segments.  Algorithmic bytes: 8*n per column.
#include "aux.cuh"

namespace zkm {

constexpr int OPN_THREADS = 256, OPN_THREAD_BITS = 5, OPN_SEG_BITS = 13;     // 256 threads x 32 coefficients
constexpr int OPN_MAX_POINTS = 3;

struct OpenPoints {
    u64 pow2[OPN_MAX_POINTS][34][2];          // z^(2^k), k = 0..33
    int npoints;
};

__device__ __forceinline__ gl2 ld_gl2(const u64 (*t)[2], int k) { return gl2(gl(t[k][0]), gl(t[k][1])); }

__global__ void __launch_bounds__(OPN_THREADS) open_segments_kernel(const u64* __restrict__ coeffs, size_t n, int log_n, OpenPoints pts,
                                                                    u64* __restrict__ partial, int nseg) {
    __shared__ u64 sh[OPN_MAX_POINTS][OPN_THREADS][2];
    const int seg = blockIdx.x, col = blockIdx.y, tid = threadIdx.x;
    const u64* c = coeffs + (size_t)col * n;
    size_t start = ((size_t)seg << OPN_SEG_BITS) + ((size_t)tid << OPN_THREAD_BITS);
    constexpr int L = 1 << OPN_THREAD_BITS;
    gl v[L];
#pragma unroll
    for (int k = 0; k < L; k++) v[k] = start + k < n ? gl(__ldg(c + start + k)) : gl::zero();
    for (int p = 0; p < pts.npoints; p++) {
        gl2 z = ld_gl2(pts.pow2[p], 0);
        gl2 acc = gl2::zero();
#pragma unroll
        for (int k = L - 1; k >= 0; k--) acc = acc * z + v[k];
        sh[p][tid][0] = acc.a.v; sh[p][tid][1] = acc.b.v;
    }
    __syncthreads();
    // power tree: level l combines neighbours 2^l apart with z^(L * 2^l)
    for (int l = 0; (1 << l) < OPN_THREADS; l++) {
        int stride = 1 << l;
        if ((tid & (2 * stride - 1)) == 0) {
            for (int p = 0; p < pts.npoints; p++) {
                gl2 lo = mk2(sh[p][tid][0], sh[p][tid][1]), hi = mk2(sh[p][tid + stride][0], sh[p][tid + stride][1]);
                gl2 r = lo + hi * ld_gl2(pts.pow2[p], OPN_THREAD_BITS + l);
                sh[p][tid][0] = r.a.v; sh[p][tid][1] = r.b.v;
            }
        }
        __syncthreads();
    }
    if (tid < pts.npoints) {
        int p = tid;
        // multiply by z^(seg * 2^SEG_BITS)
        gl2 r = mk2(sh[p][0][0], sh[p][0][1]);
        gl2 m = gl2::one();
        for (int b = 0; (seg >> b) != 0; b++)
            if ((seg >> b) & 1) m = m * ld_gl2(pts.pow2[p], OPN_SEG_BITS + b);
        r = r * m;
        size_t o = (((size_t)col * pts.npoints + p) * nseg + seg) * 2;
        partial[o] = r.a.v; partial[o + 1] = r.b.v;
    }
}

__global__ void open_sum_kernel(const u64* __restrict__ partial, int nseg, int total, u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    gl2 acc = gl2::zero();
    for (int s = 0; s < nseg; s++) acc = acc + gl2(gl(partial[((size_t)t * nseg + s) * 2]), gl(partial[((size_t)t * nseg + s) * 2 + 1]));
    out[2 * t] = acc.a.v; out[2 * t + 1] = acc.b.v;
}

void eval_polys_at_points(const u64* d_coeffs, int ncols, int log_n, const gl2* points, int npoints, u64* h_out, cudaStream_t s) {
    ZKM_CHECK(npoints >= 1 && npoints <= OPN_MAX_POINTS, "bad number of opening points");
    if (ncols == 0) return;
    size_t n = (size_t)1 << log_n;
    OpenPoints pts;
    pts.npoints = npoints;
    for (int p = 0; p < npoints; p++) {
        gl2 cur = points[p];
        for (int k = 0; k < 34; k++) { pts.pow2[p][k][0] = cur.a.v; pts.pow2[p][k][1] = cur.b.v; cur = cur * cur; }
    }
    int nseg = (int)((n + ((size_t)1 << OPN_SEG_BITS) - 1) >> OPN_SEG_BITS);
    int total = ncols * npoints;
    DevBuf partial((size_t)total * nseg * 2, s), out((size_t)total * 2, s);
    {
        ProfScope ps("openings", s, 8.0 * (double)n * ncols);
        open_segments_kernel<<<dim3(nseg, ncols), OPN_THREADS, 0, s>>>(d_coeffs, n, log_n, pts, partial.p, nseg);
        ZKM_LAUNCHED();
        open_sum_kernel<<<(total + 127) / 128, 128, 0, s>>>(partial.p, nseg, total, out.p);
        ZKM_LAUNCHED();
    }
    out.download(h_out, (size_t)total * 2);
}

}  // namespace zkm
