#include "batch.cuh"
#include <cstdlib>
#include "shard.cuh"
#include <memory>
#include <map>
#include <mutex>
#include <atomic>
#include <algorithm>

namespace zkm {

static std::unique_ptr<Ctx> g_ctx;
// A host thread may bind a worker context (zkm_b200_worker_bind): its own streams, NTT tables and arena, so that two proofs
// can be in flight on one device (the latency-bound phases of one overlap the throughput-bound kernels of the other).
static thread_local Ctx* t_ctx = nullptr;

// ---- arena (dev.cuh): one per context, because a freed block is handed to the next user on the SAME stream only.
// All arenas are registered so that a failed cudaMalloc can give back the idle blocks cached by EVERY context (with several
// proofs in flight the sibling workers' caches would otherwise stay pinned while this one runs out of memory), and the bytes
// cached device-wide are capped at half of the device memory whatever the number of contexts.
struct Arena {
    std::mutex mu;                                  // the owner thread allocates/frees; any thread may trim
    std::multimap<size_t, void*> free_blocks;       // size -> block
    size_t cached = 0, live = 0;
};
namespace {
std::mutex g_arenas_mu;
std::vector<Arena*>& all_arenas() { static std::vector<Arena*>* v = new std::vector<Arena*>(); return *v; }
std::atomic<size_t> g_cached_total{0};
size_t g_cache_limit = (size_t)64 << 30;           // set from the device size in ctx_init
Arena* arena_new() {
    Arena* a = new Arena();
    std::lock_guard<std::mutex> g(g_arenas_mu);
    all_arenas().push_back(a);
    return a;
}
void arena_unregister(Arena* a) {
    std::lock_guard<std::mutex> g(g_arenas_mu);
    auto& v = all_arenas();
    v.erase(std::remove(v.begin(), v.end(), a), v.end());
}
// heap-allocated and never destroyed: DevBufs owned by other statics may be released during process exit
Arena& g_arena0 = *arena_new();
Arena& cur_arena() { return (t_ctx && t_ctx->arena) ? *t_ctx->arena : g_arena0; }
// cached blocks are idle by construction (their last user's kernels may still be running: hence the device synchronisation)
void arena_trim_locked(Arena& a) {
    for (auto& kv : a.free_blocks) cudaFree(kv.second);
    g_cached_total -= a.cached;
    a.free_blocks.clear();
    a.cached = 0;
}
void arena_trim_one(Arena& a) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> g(a.mu);
    arena_trim_locked(a);
}
void arena_trim_all() {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> g(g_arenas_mu);
    for (Arena* a : all_arenas()) { std::lock_guard<std::mutex> ga(a->mu); arena_trim_locked(*a); }
}
}
void* arena_alloc(size_t bytes) {
    Arena& A = cur_arena();
    bytes = (bytes + 511) & ~(size_t)511;
    {
        std::lock_guard<std::mutex> g(A.mu);
        auto it = A.free_blocks.lower_bound(bytes);
        // reuse a cached block when it is not wastefully larger than the request
        if (it != A.free_blocks.end() && it->first <= bytes + bytes / 8 + 4096) {
            void* p = it->second;
            A.cached -= it->first;
            g_cached_total -= it->first;
            A.live += it->first;
            A.free_blocks.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        arena_trim_all();                          // give every context's cached blocks back and retry once
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc(") + std::to_string(bytes) + " bytes): " + cudaGetErrorString(e));
    std::lock_guard<std::mutex> g(A.mu);
    A.live += bytes;
    return p;
}
void arena_free(void* p, size_t bytes) {
    if (!p) return;
    Arena& A = cur_arena();
    bytes = (bytes + 511) & ~(size_t)511;
    if (g_cached_total.load() + bytes > g_cache_limit) arena_trim_one(A);
    std::lock_guard<std::mutex> g(A.mu);
    // blocks handed out from the cache may be larger than the request; the size recorded here is the request rounded up,
    // which is what lower_bound matched against, so re-insert under that size (never larger than the real block)
    A.live -= bytes <= A.live ? bytes : A.live;
    A.free_blocks.emplace(bytes, p);
    A.cached += bytes;
    g_cached_total += bytes;
}
void arena_trim() { arena_trim_one(cur_arena()); }
size_t arena_cached_bytes() { Arena& A = cur_arena(); std::lock_guard<std::mutex> g(A.mu); return A.cached; }

bool ctx_ready() { return (bool)g_ctx; }
Ctx& ctx() {
    if (t_ctx) return *t_ctx;
    if (!g_ctx) throw std::runtime_error("zkm_b200: not initialised (call zkm_b200_init; a CUDA device is required)");
    return *g_ctx;
}
bool blocking_sync_enabled() {
    static const bool on = std::getenv("ZKM_BLOCKING_SYNC") && atoi(std::getenv("ZKM_BLOCKING_SYNC")) != 0;
    return on;
}
cudaError_t stream_sync(cudaStream_t s) {
    if (!blocking_sync_enabled()) return cudaStreamSynchronize(s);
    static thread_local cudaEvent_t ev = nullptr;          // one per host thread, never destroyed (threads are few and long-lived)
    static thread_local int ev_device = -1;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!ev || ev_device != dev) {
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return e;
        ev_device = dev;
    }
    if ((e = cudaEventRecord(ev, s)) != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}
Ctx::~Ctx() {
    for (int k = 0; k < BOUNCE_SLOTS; k++) {
        if (bounce_free[k]) cudaEventDestroy(bounce_free[k]);
        if (bounce[k]) cudaFreeHost(bounce[k]);
    }
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
    g_cache_limit = prop.totalGlobalMem / 2;
    auto c = std::make_unique<Ctx>();
    c->device = device;
    ZKM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ZKM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    g_ctx = std::move(c);
}
Ctx* worker_create() {
    ZKM_CHECK(g_ctx, "zkm_b200: not initialised");
    ZKM_CUDA(cudaSetDevice(g_ctx->device));
    auto c = std::make_unique<Ctx>();
    c->device = g_ctx->device;
    c->arena = arena_new();
    ZKM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ZKM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    return c.release();
}
void worker_bind(Ctx* w) {
    if (w) ZKM_CUDA(cudaSetDevice(w->device));
    t_ctx = w;
}
void worker_destroy(Ctx* w) {
    if (!w) return;
    Ctx* prev = t_ctx;
    t_ctx = w;                                     // the worker's tables go back to the worker's arena
    cudaStreamSynchronize(w->stream);
    cudaStreamSynchronize(w->copy_stream);
    cudaStream_t s0 = w->stream, s1 = w->copy_stream;
    Arena* a = w->arena;
    delete w;                                      // ~NttTables releases its buffers while t_ctx still names this context
    t_ctx = (prev == w) ? nullptr : prev;
    arena_trim_one(*a);
    arena_unregister(a);
    delete a;
    cudaStreamDestroy(s0);
    cudaStreamDestroy(s1);
}
void ctx_shutdown() {
    if (!g_ctx) return;
    cudaStreamSynchronize(g_ctx->stream);
    cudaStream_t s = g_ctx->stream;
    cudaStreamSynchronize(g_ctx->copy_stream);
    cudaStreamDestroy(g_ctx->copy_stream);
    g_ctx.reset();
    arena_trim();
    cudaStreamDestroy(s);
}

// In-segment sharding applies to commitments big enough to pay for one 512-byte all-gather (tables below 2^13 rows are
// latency bound on one GPU already -- they also take the cooperative small-table quotient kernel -- and are committed in
// full by every rank).  The rule depends on the shape only, so every
// rank takes the same branch.
static bool commit_is_sharded(const Batch& b) {
    return shard().active() && b.rate_bits == 2 && (int)b.cap_height >= shard().log_segs() && b.log_n >= 13;
}
static void commit_tree(Batch& b, cudaStream_t s) {
    const Shard& sh = shard();
    merkle_alloc(b.tree, b.lde_bits(), b.cap_height, s);
    b.tree.sharded = b.sharded;
    if (b.sharded) lde_leaf_hash(b.lde.p, b.lde_n(), b.ncols, b.log_n, b.rate_bits, b.tree.digests.p, s, sh.coset_begin(), sh.coset_count(),
                                 sh.log_parts(), sh.part());
    else lde_leaf_hash(b.lde.p, b.lde_n(), b.ncols, b.log_n, b.rate_bits, b.tree.digests.p, s);
    merkle_build_from_leaf_digests(b.tree, s);
}
static void commit_lde(Batch& b) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const Shard& sh = shard();
    size_t N = b.lde_n();
    b.sharded = commit_is_sharded(b);
    b.lde.alloc((size_t)b.ncols * N, s);
    if (b.sharded) lde_coset(c.ntt, b.coeffs.p, b.n(), b.lde.p, N, b.ncols, b.log_n, b.rate_bits, s, 0, sh.coset_begin(), sh.coset_count());
    else lde_coset(c.ntt, b.coeffs.p, b.n(), b.lde.p, N, b.ncols, b.log_n, b.rate_bits, s);
    commit_tree(b, s);
}

void batch_from_coeffs_dev(Batch& b, DevBuf&& coeffs, int ncols, int log_n, int rate_bits, int cap_height) {
    ZKM_CHECK(ncols > 0, "empty polynomial batch");
    ZKM_CHECK(log_n + rate_bits >= cap_height, "cap height exceeds LDE tree height");
    ZKM_CHECK(log_n + rate_bits <= 31, "LDE too large");
    b.ncols = ncols; b.log_n = log_n; b.rate_bits = rate_bits; b.cap_height = cap_height;
    b.coeffs = std::move(coeffs);
    commit_lde(b);
}

void batch_from_values_grouped_dev(Batch& b, const u64* values, DevBuf&& coeffs, int ncols, int log_n, int rate_bits, int cap_height,
                                   const std::vector<int>& col_ends, const std::function<void(size_t)>& wait_group) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    ZKM_CHECK(ncols > 0 && !col_ends.empty() && col_ends.back() == ncols, "bad column groups");
    ZKM_CHECK(log_n + rate_bits >= cap_height, "cap height exceeds LDE tree height");
    ZKM_CHECK(log_n + rate_bits <= 31, "LDE too large");
    b.ncols = ncols; b.log_n = log_n; b.rate_bits = rate_bits; b.cap_height = cap_height;
    b.coeffs = std::move(coeffs);
    size_t n = b.n(), N = b.lde_n();
    b.sharded = commit_is_sharded(b);
    b.lde.alloc((size_t)ncols * N, s);
    int c0 = 0;
    for (size_t k = 0; k < col_ends.size(); k++) {
        int c1 = col_ends[k];
        ZKM_CHECK(c1 > c0, "bad column groups");
        wait_group(k);
        ntt_inverse(c.ntt, values + (size_t)c0 * n, n, b.coeffs.p + (size_t)c0 * n, n, c1 - c0, log_n, s);
        if (b.sharded) lde_coset(c.ntt, b.coeffs.p + (size_t)c0 * n, n, b.lde.p + (size_t)c0 * N, N, c1 - c0, log_n, rate_bits, s, 0,
                                 shard().coset_begin(), shard().coset_count());
        else lde_coset(c.ntt, b.coeffs.p + (size_t)c0 * n, n, b.lde.p + (size_t)c0 * N, N, c1 - c0, log_n, rate_bits, s);
        c0 = c1;
    }
    commit_tree(b, s);
}

void batch_from_values_dev(Batch& b, DevBuf&& values, int ncols, int log_n, int rate_bits, int cap_height) {
    Ctx& c = ctx();
    ZKM_CHECK(ncols > 0, "empty polynomial batch");
    size_t n = (size_t)1 << log_n;
    ntt_inverse(c.ntt, values.p, n, values.p, n, ncols, log_n, c.stream);
    batch_from_coeffs_dev(b, std::move(values), ncols, log_n, rate_bits, cap_height);
}

}  // namespace zkm
