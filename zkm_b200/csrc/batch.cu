#include "batch.cuh"
#include <memory>

namespace zkm {

static std::unique_ptr<Ctx> g_ctx;

bool ctx_ready() { return (bool)g_ctx; }
Ctx& ctx() {
    if (!g_ctx) throw std::runtime_error("zkm_b200: not initialised (call zkm_b200_init; a CUDA device is required)");
    return *g_ctx;
}
void ctx_init(int device) {
    if (g_ctx) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw std::runtime_error(std::string("zkm_b200: no CUDA device available (") + cudaGetErrorString(e) +
                                 "); this library has no CPU fallback");
    ZKM_CHECK(device >= 0 && device < count, "zkm_b200: device index out of range");
    ZKM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ZKM_CUDA(cudaGetDeviceProperties(&prop, device));
    ZKM_CHECK(prop.major >= 10, "zkm_b200: kernels are built for sm_100a (Blackwell) only");
    auto c = std::make_unique<Ctx>();
    c->device = device;
    ZKM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // keep freed blocks in the pool: the prover allocates/free multi-GB buffers per table
    cudaMemPool_t pool;
    ZKM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thresh = ~0ull;
    ZKM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    g_ctx = std::move(c);
}
void ctx_shutdown() {
    if (!g_ctx) return;
    cudaStreamSynchronize(g_ctx->stream);
    cudaStream_t s = g_ctx->stream;
    g_ctx.reset();
    cudaStreamDestroy(s);
}

static void commit_lde(Batch& b) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    size_t N = b.lde_n();
    b.lde.alloc((size_t)b.ncols * N, s);
    lde_coset(c.ntt, b.coeffs.p, b.n(), b.lde.p, N, b.ncols, b.log_n, b.rate_bits, s);
    merkle_alloc(b.tree, b.lde_bits(), b.cap_height, s);
    lde_leaf_hash(b.lde.p, N, b.ncols, b.log_n, b.rate_bits, b.tree.digests.p, s);
    merkle_build_from_leaf_digests(b.tree, s);
}

void batch_from_coeffs_dev(Batch& b, DevBuf&& coeffs, int ncols, int log_n, int rate_bits, int cap_height) {
    ZKM_CHECK(ncols > 0, "empty polynomial batch");
    ZKM_CHECK(log_n + rate_bits >= cap_height, "cap height exceeds LDE tree height");
    ZKM_CHECK(log_n + rate_bits <= 31, "LDE too large");
    b.ncols = ncols; b.log_n = log_n; b.rate_bits = rate_bits; b.cap_height = cap_height;
    b.coeffs = std::move(coeffs);
    commit_lde(b);
}

void batch_from_values_dev(Batch& b, DevBuf&& values, int ncols, int log_n, int rate_bits, int cap_height) {
    Ctx& c = ctx();
    ZKM_CHECK(ncols > 0, "empty polynomial batch");
    size_t n = (size_t)1 << log_n;
    ntt_inverse(c.ntt, values.p, n, values.p, n, ncols, log_n, c.stream);
    batch_from_coeffs_dev(b, std::move(values), ncols, log_n, rate_bits, cap_height);
}

}  // namespace zkm
