// C ABI (include/zkm_b200.h).  Every entry point catches C++ exceptions and reports them through
// the (status, char** err) convention of the reference's own FFI (recursion/src/snark/snarks.rs:7-20).
#include "../../include/zkm_b200.h"
#include "batch.cuh"
#include "poseidon_v2.cuh"
#include "poseidon_host.h"
#include "prover.cuh"
#include "shard.cuh"
#include "tables/systems.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <cstdio>

using namespace zkm;

struct zkm_batch { Batch b; };

static int fail(char** err, const std::exception& e) {
    if (err) {
        const char* w = e.what();
        size_t n = strlen(w);
        char* m = (char*)malloc(n + 1);
        if (m) memcpy(m, w, n + 1);
        *err = m;
    }
    return -1;
}
#define ZKM_API_BEGIN if (err) *err = nullptr; try {
#define ZKM_API_END } catch (const std::exception& e) { return fail(err, e); } return 0;

extern "C" {

void zkm_b200_free_string(char* s) { free(s); }
void zkm_b200_free(void* p) { free(p); }

void zkm_b200_standard_fast_config(zkm_stark_config_t* c) {
    c->rate_bits = 2; c->cap_height = 4; c->pow_bits = 16; c->num_queries = 37;
    c->num_challenges = 2; c->arity_bits = 4; c->final_poly_bits = 5;
}

int zkm_b200_init(int device, char** err) {
    ZKM_API_BEGIN
    ctx_init(device);
    ZKM_API_END
}
int zkm_b200_shutdown(char** err) {
    ZKM_API_BEGIN
    ctx_shutdown();
    ZKM_API_END
}
uint64_t zkm_b200_launch_count(void) { return g_launch_count; }
int zkm_b200_sync(char** err) {
    ZKM_API_BEGIN
    ZKM_CUDA(stream_sync(ctx().stream));
    ZKM_API_END
}

int zkm_b200_worker_create(zkm_worker_t** out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(out, "null argument");
    *out = (zkm_worker_t*)worker_create();
    ZKM_API_END
}
int zkm_b200_worker_bind(zkm_worker_t* w, char** err) {
    ZKM_API_BEGIN
    worker_bind((Ctx*)w);
    ZKM_API_END
}
void zkm_b200_worker_destroy(zkm_worker_t* w) {
    try { worker_destroy((Ctx*)w); } catch (...) {}
}

int zkm_b200_shard_unique_id(uint8_t out[128], char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(out, "null argument");
    shard_unique_id(out);
    ZKM_API_END
}
int zkm_b200_shard_init(int rank, int world, const uint8_t id[128], char** err) {
    ZKM_API_BEGIN
    shard_init(rank, world, id);
    ZKM_API_END
}
int zkm_b200_shard_shutdown(char** err) {
    ZKM_API_BEGIN
    shard_shutdown();
    ZKM_API_END
}

static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
int zkm_b200_timer_start(char** err) {
    ZKM_API_BEGIN
    if (!g_t0) { ZKM_CUDA(cudaEventCreate(&g_t0)); ZKM_CUDA(cudaEventCreate(&g_t1)); }
    ZKM_CUDA(cudaEventRecord(g_t0, ctx().stream));
    ZKM_API_END
}
int zkm_b200_timer_stop(double* ms, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(g_t0, "timer not started");
    ZKM_CUDA(cudaEventRecord(g_t1, ctx().stream));
    ZKM_CUDA(cudaEventSynchronize(g_t1));
    float f = 0;
    ZKM_CUDA(cudaEventElapsedTime(&f, g_t0, g_t1));
    *ms = f;
    ZKM_API_END
}
void zkm_b200_profile_enable(int on) { prof_enable(on != 0); }
int zkm_b200_profile_reset(char** err) {
    ZKM_API_BEGIN
    prof_reset();
    ZKM_API_END
}
int zkm_b200_profile_get(const char* family, double* ms, uint64_t* launches, double* bytes, char** err) {
    ZKM_API_BEGIN
    unsigned long long l = 0; double m = 0, b = 0;
    bool ok = prof_get(family, &m, &l, &b);
    if (ms) *ms = m;
    if (launches) *launches = l;
    if (bytes) *bytes = b;
    if (!ok) throw std::runtime_error(std::string("no such kernel family: ") + family);
    ZKM_API_END
}
void zkm_b200_timing_enable(int on) { scopes_enable(on != 0); }
char* zkm_b200_last_timing(void) {
    const std::string& t = scopes_last();
    char* m = (char*)malloc(t.size() + 1);
    if (m) memcpy(m, t.c_str(), t.size() + 1);
    return m;
}
int zkm_b200_profile_get_traffic(const char* family, double* aux, char** err) {
    ZKM_API_BEGIN
    double a = 0;
    if (!prof_get(family, nullptr, nullptr, nullptr, &a)) throw std::runtime_error(std::string("no such kernel family: ") + family);
    if (aux) *aux = a;
    ZKM_API_END
}
char* zkm_b200_profile_families(void) {
    std::string n;
    try { n = prof_names(); } catch (...) {}
    char* m = (char*)malloc(n.size() + 1);
    if (m) memcpy(m, n.c_str(), n.size() + 1);
    return m;
}

static DevBuf upload_table(const zkm_table_t* t) {
    Ctx& c = ctx();
    ZKM_CHECK(t && t->cols && t->ncols > 0, "null/empty table");
    size_t n = (size_t)1 << t->log_n;
    DevBuf d((size_t)t->ncols * n, c.stream);
    for (uint32_t i = 0; i < t->ncols; i++) {
        ZKM_CHECK(t->cols[i] != nullptr, "null column pointer");
        d.upload(t->cols[i], n, (size_t)i * n);
    }
    return d;
}

int zkm_b200_commit_values(const zkm_table_t* table, uint32_t rate_bits, uint32_t cap_height, zkm_batch_t** out,
                           uint64_t* cap_out, char** err) {
    ZKM_API_BEGIN
    DevBuf d = upload_table(table);
    auto* h = new zkm_batch;
    try {
        batch_from_values_dev(h->b, std::move(d), table->ncols, table->log_n, rate_bits, cap_height);
    } catch (...) { delete h; throw; }
    if (cap_out) memcpy(cap_out, h->b.tree.cap.data(), h->b.tree.cap.size() * sizeof(u64));
    *out = h;
    ZKM_API_END
}

int zkm_b200_commit_coeffs(const zkm_table_t* table, uint32_t rate_bits, uint32_t cap_height, zkm_batch_t** out,
                           uint64_t* cap_out, char** err) {
    ZKM_API_BEGIN
    DevBuf d = upload_table(table);
    auto* h = new zkm_batch;
    try {
        batch_from_coeffs_dev(h->b, std::move(d), table->ncols, table->log_n, rate_bits, cap_height);
    } catch (...) { delete h; throw; }
    if (cap_out) memcpy(cap_out, h->b.tree.cap.data(), h->b.tree.cap.size() * sizeof(u64));
    *out = h;
    ZKM_API_END
}

int zkm_b200_commit_values_device(const uint64_t* d_values, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
                                  uint32_t cap_height, zkm_batch_t** out, uint64_t* cap_out, char** err) {
    ZKM_API_BEGIN
    Ctx& c = ctx();
    size_t n = (size_t)1 << log_n;
    DevBuf d((size_t)ncols * n, c.stream);
    ZKM_CUDA(cudaMemcpyAsync(d.p, d_values, (size_t)ncols * n * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream));
    auto* h = new zkm_batch;
    try {
        batch_from_values_dev(h->b, std::move(d), ncols, log_n, rate_bits, cap_height);
    } catch (...) { delete h; throw; }
    if (cap_out) memcpy(cap_out, h->b.tree.cap.data(), h->b.tree.cap.size() * sizeof(u64));
    *out = h;
    ZKM_API_END
}

// One table built on the device from its operation log; returns the table height.
static size_t table_from_ops_dev(int table, const u64* ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    switch (table) {
        case tables::T_MEMORY: return memory_generate_trace_dev(ops, n_ops, cols, s);
        case tables::T_ARITHMETIC: return arithmetic_generate_trace_dev(ops, n_ops, cols, s);
        case tables::T_LOGIC: return logic_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_POSEIDON: return poseidon_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_KECCAK: return keccak_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_POSEIDON_SPONGE: return poseidon_sponge_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_KECCAK_SPONGE: return keccak_sponge_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_SHA_EXTEND: return sha_extend_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_SHA_EXTEND_SPONGE: return sha_extend_sponge_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_SHA_COMPRESS: return sha_compress_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        case tables::T_SHA_COMPRESS_SPONGE: return sha_compress_sponge_generate_trace_dev(ops, n_ops, min_rows, cols, s);
        default: throw std::runtime_error(std::string("no device-side generator for table ") + tables::table_name(table) +
                                          " (the Cpu table is the interpreter's own rows: pass them row-major)");
    }
}

// Host -> device copy of `bytes` from `src` on stream `cs`.  Pinned (or registered) sources go straight to cudaMemcpyAsync.
// Pageable sources -- what a `Vec<F>` column of the reference is -- go through the context's pinned bounce ring: the chunk is
// copied into a free slot by up to 4 host threads while the DMA engine is still draining the previous slots (the driver's own
// pageable path stages and copies serially on the calling thread: 7.4 GB/s end to end on the bench host against 9.2 needed).
// ZKM_BOUNCE=0 switches the ring off (A/B).
static bool source_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static cudaError_t upload_bounced(Ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t cs) {
    static const bool enabled = !(std::getenv("ZKM_BOUNCE") && atoi(std::getenv("ZKM_BOUNCE")) == 0);
    if (!enabled || bytes < ((size_t)1 << 20) || !source_is_pageable(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs);
    for (size_t off = 0; off < bytes;) {
        const size_t len = std::min(bytes - off, Ctx::BOUNCE_BYTES);
        const int k = c->bounce_next;
        c->bounce_next = (k + 1) % Ctx::BOUNCE_SLOTS;
        cudaError_t e;
        if (!c->bounce[k]) {
            if ((e = cudaHostAlloc(&c->bounce[k], Ctx::BOUNCE_BYTES, cudaHostAllocDefault)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&c->bounce_free[k], cudaEventDisableTiming | (blocking_sync_enabled() ? cudaEventBlockingSync : 0))) != cudaSuccess) return e;
        } else if ((e = cudaEventSynchronize(c->bounce_free[k])) != cudaSuccess) return e;      // the slot's last DMA has finished
        const char* from = (const char*)src + off;
        char* slot = (char*)c->bounce[k];
        const int parts = len >= ((size_t)4 << 20) ? 4 : 1;
        const size_t per = (len / parts + 63) & ~(size_t)63;
        std::thread helpers[3];
        for (int t = 1; t < parts; t++) {
            const size_t b = std::min(len, (size_t)t * per), e2 = std::min(len, (size_t)(t + 1) * per);
            helpers[t - 1] = std::thread([=] { memcpy(slot + b, from + b, e2 - b); });
        }
        memcpy(slot, from, std::min(len, per));
        for (int t = 1; t < parts; t++) helpers[t - 1].join();
        if ((e = cudaMemcpyAsync((char*)dst + off, slot, len, cudaMemcpyHostToDevice, cs)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(c->bounce_free[k], cs)) != cudaSuccess) return e;
        off += len;
    }
    return cudaSuccess;
}

static int prove_common(int system_id, const zkm_table_t* tables_in, uint32_t num_tables, const uint64_t* const* d_tables,
                        const uint32_t* roots_before, const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                        const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words,
                        const zkm_table_rows_t* row_tables = nullptr, const zkm_op_log_t* op_logs = nullptr) {
    Ctx& c = ctx();
    ZKM_CHECK(tables_in && roots_before && roots_after && cfg && proof_out && proof_words, "null argument");
    // tables given as rows (zkm_b200_prove_with_trace_rows): shape from the row descriptor, data uploaded as one block
    std::vector<zkm_table_t> merged(tables_in, tables_in + num_tables);
    for (uint32_t t = 0; t < num_tables && row_tables; t++)
        if (row_tables[t].rows) { merged[t].cols = nullptr; merged[t].ncols = row_tables[t].ncols; merged[t].log_n = row_tables[t].log_n; }
    // tables generated on the device from their operation logs (zkm_b200_prove_with_ops): they never leave HBM
    std::vector<char> generated(num_tables, 0);
    std::vector<DevBuf> generated_bufs(num_tables);
    if (op_logs) {
        ZKM_CHECK(system_id == tables::SYSTEM_ALL_STARK && !d_tables, "operation logs are taken by the AllStark host entry points only");
        // min_rows of Traces::into_tables (witness/traces.rs:239-240): max(number of cap elements, MIN_TRACE_LEN = 64)
        const size_t min_rows = std::max<size_t>((size_t)1 << cfg->cap_height, 64);
        for (uint32_t t = 0; t < num_tables; t++) {
            if (!op_logs[t].ops) continue;
            const size_t n = table_from_ops_dev((int)t, op_logs[t].ops, op_logs[t].n_ops, min_rows, generated_bufs[t], c.stream);
            uint32_t lg = 0;
            while (((size_t)1 << lg) < n) lg++;
            merged[t].cols = nullptr; merged[t].ncols = (uint32_t)tables::table_num_columns((int)t); merged[t].log_n = lg;
            generated[t] = 1;
        }
    }
    const zkm_table_t* tables = merged.data();
    std::vector<DevBuf> row_staging(num_tables);
    DevBuf bad_flag(1, c.stream);              // raised by the device-side Arithmetic range check (rows path)
    ZKM_CHECK(userdata || userdata_len == 0, "null userdata");
    StarkCfg sc;
    sc.rate_bits = cfg->rate_bits; sc.cap_height = cfg->cap_height; sc.pow_bits = cfg->pow_bits; sc.num_queries = cfg->num_queries;
    sc.num_challenges = cfg->num_challenges; sc.arity_bits = cfg->arity_bits; sc.final_poly_bits = cfg->final_poly_bits;
    std::vector<TableInput> in(num_tables);
    auto T0 = std::chrono::steady_clock::now();
    // declared before the per-table loop below so that an exception thrown there still destroys the events created so far
    struct Uploader {
        std::thread th; std::mutex mu; std::condition_variable cv; std::vector<char> done; std::vector<int> groups_done;
        std::string error;
        ~Uploader() { if (th.joinable()) th.join(); }
    } up;
    struct EventGuard {
        std::vector<TableInput>& v; cudaStream_t cs; Uploader& u;
        ~EventGuard() { if (u.th.joinable()) u.th.join(); cudaStreamSynchronize(cs); for (auto& x : v) { if (x.ready) cudaEventDestroy(x.ready); for (cudaEvent_t e : x.group_ready) if (e) cudaEventDestroy(e); } }
    } guard{in, c.copy_stream, up};
    for (uint32_t t = 0; t < num_tables; t++) {
        in[t].ncols = tables[t].ncols; in[t].log_n = tables[t].log_n;
        ZKM_CHECK(tables[t].log_n <= 26, "trace too long");
        size_t n = (size_t)1 << tables[t].log_n;
        if (d_tables) {
            ZKM_CHECK(d_tables[t] != nullptr, "null device table");
            in[t].values.alloc((size_t)tables[t].ncols * n, c.stream);
            ZKM_CUDA(cudaMemcpyAsync(in[t].values.p, d_tables[t], (size_t)tables[t].ncols * n * sizeof(u64), cudaMemcpyDeviceToDevice, c.stream));
        } else {
            // uploads run on the copy stream from a helper thread (cudaMemcpyAsync blocks the calling thread for
            // pageable and, on this platform, also for pinned sources), so table t+1.. stream in while table t is
            // being committed; the prover waits on `ready` before touching a buffer
            const zkm_table_t* tb = &tables[t];
            if (generated[t]) { in[t].values = std::move(generated_bufs[t]); continue; }     // already on the compute stream
            const bool as_rows = row_tables && row_tables[t].rows;
            ZKM_CHECK((tb->cols || as_rows) && tb->ncols > 0, "null/empty table");
            if (!as_rows) for (uint32_t i = 0; i < tb->ncols; i++) ZKM_CHECK(tb->cols[i] != nullptr, "null column pointer");
            in[t].values.alloc((size_t)tb->ncols * n, c.stream);
            ZKM_CUDA(cudaEventCreateWithFlags(&in[t].ready, cudaEventDisableTiming));
            if (as_rows) {
                ZKM_CHECK(tb->log_n >= 5, "row-major tables need at least 32 rows");
                row_staging[t].alloc((size_t)tb->ncols * n, c.stream);
                continue;                       // one block: no column groups
            }
            // tables above 256 MB arrive in 4 column groups so that their NTTs start before the whole table is resident
            size_t bytes = (size_t)tb->ncols * n * sizeof(u64);
            size_t group_threshold = (size_t)256 << 20;
            if (const char* gt = std::getenv("ZKM_GROUP_BYTES")) group_threshold = (size_t)strtoull(gt, nullptr, 10);   // tests
            if (bytes > group_threshold && tb->ncols >= 8) {
                for (int g = 1; g <= 4; g++) in[t].group_ends.push_back((int)((size_t)tb->ncols * g / 4));
                in[t].group_ready.resize(4, nullptr);
                for (int g = 0; g < 4; g++) ZKM_CUDA(cudaEventCreateWithFlags(&in[t].group_ready[g], cudaEventDisableTiming));
            }
        }
    }
    if (!d_tables) {
        int device = c.device;
        cudaStream_t cs = c.copy_stream;
        std::vector<u64*> dst(num_tables);
        for (uint32_t t = 0; t < num_tables; t++) dst[t] = in[t].values.p;
        std::vector<cudaEvent_t> evs(num_tables);
        for (uint32_t t = 0; t < num_tables; t++) evs[t] = in[t].ready;
        // smallest table first (the prover commits in the same order, prover.cu commit_order): the short commits run while the
        // long uploads are still in flight, so the 2 GB CPU table never stalls the device (it did for ~15 ms in table order)
        std::vector<size_t> bytes(num_tables);
        for (uint32_t t = 0; t < num_tables; t++) bytes[t] = (size_t)tables[t].ncols << tables[t].log_n;
        std::vector<size_t> order = commit_order(bytes);
        up.done.assign(num_tables, 0);
        for (uint32_t t = 0; t < num_tables; t++) up.done[t] = generated[t];
        up.groups_done.assign(num_tables, 0);
        std::vector<std::vector<int>> gends(num_tables);
        std::vector<std::vector<cudaEvent_t>> gevs(num_tables);
        for (uint32_t t = 0; t < num_tables; t++) { gends[t] = in[t].group_ends; gevs[t] = in[t].group_ready; }
        std::vector<const u64*> rows_src(num_tables, nullptr);
        std::vector<u64*> rows_stage(num_tables, nullptr);
        for (uint32_t t = 0; t < num_tables; t++)
            if (row_tables && row_tables[t].rows) { rows_src[t] = row_tables[t].rows; rows_stage[t] = row_staging[t].p; }
        // the staging buffers and `values` were allocated on the compute stream: the copy stream must not run ahead of that
        cudaEvent_t alloc_done;
        ZKM_CUDA(cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming));
        ZKM_CUDA(cudaEventRecord(alloc_done, c.stream));
        ZKM_CUDA(cudaStreamWaitEvent(cs, alloc_done, 0));
        ZKM_CUDA(cudaEventDestroy(alloc_done));
        // AllStark's Arithmetic table handed over as rows comes without its range-check columns (arithmetic_stark.rs:155-192
        // builds them after the transposition, :127-153): they are generated on the device
        const int arith_rows_table = (system_id == tables::SYSTEM_ALL_STARK && rows_src[tables::T_ARITHMETIC]) ? tables::T_ARITHMETIC : -1;
        unsigned* d_bad = (unsigned*)bad_flag.p;
        Ctx* cptr = &c;
        up.th = std::thread([&up, cptr, tables, num_tables, device, cs, dst, evs, order, gends, gevs, rows_src, rows_stage, arith_rows_table, d_bad] {
            cudaSetDevice(device);
            for (size_t t : order) {
                if (!evs[t]) continue;                      // generated on the device: nothing to upload
                size_t n = (size_t)1 << tables[t].log_n;
                cudaError_t e = cudaSuccess;
                size_t g = 0;
                if (rows_src[t]) {
                    // rows as generated (one contiguous block), transposed into column-major on the device
                    e = upload_bounced(cptr, rows_stage[t], rows_src[t], (size_t)tables[t].ncols * n * sizeof(u64), cs);
                    if (e == cudaSuccess) {
                        try {
                            transpose_rows_to_cols(rows_stage[t], dst[t], n, (int)tables[t].ncols, cs);
                            if ((int)t == arith_rows_table) {
                                // rows straight from Operation::to_rows: the range-check columns are generated here
                                namespace ar = tables::arithmetic;
                                arith_generate_range_checks(dst[t], n, ar::START_SHARED_COLS, ar::NUM_SHARED_COLS, ar::RANGE_COUNTER,
                                                            ar::RC_FREQUENCIES, d_bad, cs);
                                unsigned bad = 0;
                                e = cudaMemcpyAsync(&bad, d_bad, sizeof(unsigned), cudaMemcpyDeviceToHost, cs);
                                if (e == cudaSuccess) e = stream_sync(cs);
                                if (e == cudaSuccess && bad) {
                                    std::lock_guard<std::mutex> lk(up.mu);
                                    if (up.error.empty()) up.error = "column value exceeds the max range value 65536";
                                }
                            }
                        } catch (const std::exception& ex) {
                            std::lock_guard<std::mutex> lk(up.mu);
                            if (up.error.empty()) up.error = ex.what();
                        }
                    }
                } else
                // small tables (Keccak: 2431 columns of 512 bytes) would cost one driver call per column (~3 us each, 10 ms in
                // total for the 8 small tables of a segment): gather them on the host and upload each with a single copy
                if ((size_t)tables[t].ncols * n * sizeof(u64) <= ((size_t)8 << 20) && tables[t].ncols > 1) {
                    std::vector<u64> pack((size_t)tables[t].ncols * n);
                    for (uint32_t i = 0; i < tables[t].ncols; i++) memcpy(pack.data() + (size_t)i * n, tables[t].cols[i], n * sizeof(u64));
                    e = cudaMemcpyAsync(dst[t], pack.data(), pack.size() * sizeof(u64), cudaMemcpyHostToDevice, cs);
                    if (e == cudaSuccess) e = stream_sync(cs);      // `pack` is released at the end of this scope
                } else {
                // Large tables of concurrent proofs (worker contexts) take turns on the PCIe link instead of splitting it: a
                // proof whose upload owns the link has its data after 1/k of the time k interleaved uploads would need and
                // starts computing while the next proof uploads (ZKM_UPLOAD_FIFO=0: A/B switch).
                static std::mutex link_mu;
                static const bool fifo = !(std::getenv("ZKM_UPLOAD_FIFO") && atoi(std::getenv("ZKM_UPLOAD_FIFO")) == 0);
                std::unique_lock<std::mutex> link(link_mu, std::defer_lock);
                // (pageable sources are bound by the host-side copy into the bounce ring, not by the link: they stay concurrent)
                if (fifo && (size_t)tables[t].ncols * n * sizeof(u64) >= ((size_t)64 << 20) && !source_is_pageable(tables[t].cols[0])) link.lock();
                for (uint32_t i = 0; i < tables[t].ncols && e == cudaSuccess; i++) {
                    e = upload_bounced(cptr, dst[t] + (size_t)i * n, tables[t].cols[i], n * sizeof(u64), cs);
                    if (e == cudaSuccess && g < gends[t].size() && (int)i + 1 == gends[t][g]) {
                        e = cudaEventRecord(gevs[t][g], cs);
                        g++;
                        std::lock_guard<std::mutex> lk(up.mu);
                        up.groups_done[t] = (int)g;
                        up.cv.notify_all();
                    }
                }
                if (link.owns_lock() && e == cudaSuccess) e = stream_sync(cs);       // the link is free once this table has landed
                }
                if (e == cudaSuccess) e = cudaEventRecord(evs[t], cs);
                std::lock_guard<std::mutex> lk2(up.mu);
                if (e != cudaSuccess && up.error.empty()) up.error = cudaGetErrorString(e);
                up.done[t] = 1;
                up.cv.notify_all();
            }
        });
        for (uint32_t t = 0; t < num_tables; t++)
            in[t].wait_group = [&up, t](size_t k) {
                std::unique_lock<std::mutex> lk(up.mu);
                up.cv.wait(lk, [&] { return up.groups_done[t] > (int)k || up.done[t] != 0; });
                if (!up.error.empty()) throw CudaError("trace upload failed: " + up.error);
            };
        for (uint32_t t = 0; t < num_tables; t++)
            in[t].wait_recorded = [&up, t] {
                std::unique_lock<std::mutex> lk(up.mu);
                up.cv.wait(lk, [&] { return up.done[t] != 0; });
                if (!up.error.empty()) throw CudaError("trace upload failed: " + up.error);
            };
    }
    PublicInputs pv;
    for (int i = 0; i < 8; i++) { pv.roots_before[i] = roots_before[i]; pv.roots_after[i] = roots_after[i]; }
    pv.userdata.assign(userdata, userdata + userdata_len);
    if (std::getenv("ZKM_TRACE")) {
        fprintf(stderr, "[zkm_b200] uploads issued  %9.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count());
        cudaStreamSynchronize(c.stream);
        fprintf(stderr, "[zkm_b200] inputs resident %9.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count());
    }
    std::vector<u64> w = prove_system(system_id, sc, in, pv);
    if (std::getenv("ZKM_TRACE"))
        fprintf(stderr, "[zkm_b200] prove total     %9.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count());
    uint64_t* out = (uint64_t*)malloc(w.size() * sizeof(u64));
    ZKM_CHECK(out, "out of host memory");
    memcpy(out, w.data(), w.size() * sizeof(u64));
    *proof_out = out; *proof_words = w.size();
    return 0;
}

int zkm_b200_prove_system(int system_id, const zkm_table_t* tables, uint32_t num_tables, const uint32_t* roots_before,
                          const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len, const zkm_stark_config_t* cfg,
                          uint64_t** proof_out, size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    prove_common(system_id, tables, num_tables, nullptr, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words);
    ZKM_API_END
}

int zkm_b200_prove_system_device(int system_id, const zkm_table_t* shapes, const uint64_t* const* d_tables, uint32_t num_tables,
                                 const uint32_t* roots_before, const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                                 const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(d_tables, "null device tables");
    prove_common(system_id, shapes, num_tables, d_tables, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words);
    ZKM_API_END
}

int zkm_b200_prove_with_traces(const zkm_table_t tables[12], const uint32_t roots_before[8], const uint32_t roots_after[8],
                               const uint8_t* userdata, uint32_t userdata_len, const zkm_stark_config_t* cfg, uint64_t** proof_out,
                               size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    prove_common(tables::SYSTEM_ALL_STARK, tables, 12, nullptr, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words);
    ZKM_API_END
}
int zkm_b200_prove_with_memory_ops(const zkm_table_t* tables, const zkm_table_rows_t* row_tables, const uint64_t* memory_ops,
                                   size_t n_memory_ops, const uint32_t* roots_before, const uint32_t* roots_after,
                                   const uint8_t* userdata, uint32_t userdata_len, const zkm_stark_config_t* cfg,
                                   uint64_t** proof_out, size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(memory_ops, "null argument");
    zkm_op_log_t logs[12] = {};
    logs[tables::T_MEMORY].ops = memory_ops; logs[tables::T_MEMORY].n_ops = n_memory_ops;
    prove_common(tables::SYSTEM_ALL_STARK, tables, 12, nullptr, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words,
                 row_tables, logs);
    ZKM_API_END
}
int zkm_b200_prove_with_ops(const zkm_table_t* tables, const zkm_table_rows_t* row_tables, const zkm_op_log_t* op_logs,
                            const uint32_t* roots_before, const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                            const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(op_logs, "null argument");
    prove_common(tables::SYSTEM_ALL_STARK, tables, 12, nullptr, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words,
                 row_tables, op_logs);
    ZKM_API_END
}
int zkm_b200_table_from_ops(uint32_t table, const uint64_t* ops, size_t n_ops, uint32_t min_rows, uint64_t** cols_out, uint32_t* log_n_out,
                            char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK((ops || n_ops == 0) && cols_out && log_n_out, "null argument");
    Ctx& c = ctx();
    DevBuf cols;
    const size_t n = table_from_ops_dev((int)table, ops, n_ops, min_rows, cols, c.stream);
    const size_t words = (size_t)tables::table_num_columns((int)table) * n;
    uint64_t* out = (uint64_t*)malloc(words * sizeof(u64));
    ZKM_CHECK(out, "out of host memory");
    cols.download(out, words);
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    *cols_out = out; *log_n_out = lg;
    ZKM_API_END
}

int zkm_b200_stage_table(int system_id, uint32_t table_index, const zkm_table_t* table, const zkm_stark_config_t* cfg,
                         const uint64_t* ctl_challenges, const uint64_t* alphas, const uint64_t* zeta, uint64_t** aux_out, uint32_t* num_aux_out,
                         uint64_t** quotient_out, uint64_t** openings_out, size_t* openings_words, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(table && table->cols && cfg && ctl_challenges && alphas && zeta && aux_out && num_aux_out && quotient_out && openings_out &&
              openings_words, "null argument");
    ZKM_CHECK(table->log_n <= 24 && table->ncols > 0, "bad table shape");
    Ctx& c = ctx();
    StarkCfg sc;
    sc.rate_bits = cfg->rate_bits; sc.cap_height = cfg->cap_height; sc.pow_bits = cfg->pow_bits; sc.num_queries = cfg->num_queries;
    sc.num_challenges = cfg->num_challenges; sc.arity_bits = cfg->arity_bits; sc.final_poly_bits = cfg->final_poly_bits;
    ZKM_CHECK(sc.num_challenges >= 1 && sc.num_challenges <= (unsigned)MAX_CHALLENGES, "unsupported num_challenges");
    AuxChallenges ch = {};
    ch.count = (int)sc.num_challenges;
    for (unsigned k = 0; k < sc.num_challenges; k++) {
        ZKM_CHECK(ctl_challenges[2 * k] < GL_P && ctl_challenges[2 * k + 1] < GL_P && alphas[k] < GL_P, "challenge is not a canonical field element");
        ch.beta[k] = ctl_challenges[2 * k]; ch.gamma[k] = ctl_challenges[2 * k + 1];
    }
    ZKM_CHECK(zeta[0] < GL_P && zeta[1] < GL_P, "challenge is not a canonical field element");
    const size_t n = (size_t)1 << table->log_n;
    DevBuf values((size_t)table->ncols * n, c.stream);
    for (uint32_t i = 0; i < table->ncols; i++) {
        ZKM_CHECK(table->cols[i], "null column pointer");
        values.upload(table->cols[i], n, (size_t)i * n);
    }
    std::vector<u64> aux, quot, open;
    stage_single_table(system_id, (int)table_index, sc, std::move(values), (int)table->ncols, (int)table->log_n, ch, alphas,
                       gl2(gl(zeta[0]), gl(zeta[1])), aux, quot, open);
    auto dup = [](const std::vector<u64>& v) {
        uint64_t* p = (uint64_t*)malloc(std::max<size_t>(1, v.size()) * sizeof(u64));
        ZKM_CHECK(p, "out of host memory");
        memcpy(p, v.data(), v.size() * sizeof(u64));
        return p;
    };
    *aux_out = dup(aux); *num_aux_out = (uint32_t)(aux.size() / n);
    *quotient_out = dup(quot); *openings_out = dup(open); *openings_words = open.size();
    ZKM_API_END
}

int zkm_b200_memory_trace(const uint64_t* ops, size_t n_ops, uint64_t** cols_out, uint32_t* log_n_out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(ops && cols_out && log_n_out, "null argument");
    Ctx& c = ctx();
    DevBuf cols;
    size_t n = memory_generate_trace_dev(ops, n_ops, cols, c.stream);
    uint64_t* out = (uint64_t*)malloc(13 * n * sizeof(u64));
    ZKM_CHECK(out, "out of host memory");
    cols.download(out, 13 * n);
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    *cols_out = out; *log_n_out = lg;
    ZKM_API_END
}

int zkm_b200_prove_with_trace_rows(const zkm_table_t* tables, const zkm_table_rows_t* row_tables, const uint32_t* roots_before,
                                   const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                                   const zkm_stark_config_t* cfg, uint64_t** proof_out, size_t* proof_words, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(row_tables, "null argument");
    prove_common(tables::SYSTEM_ALL_STARK, tables, 12, nullptr, roots_before, roots_after, userdata, userdata_len, cfg, proof_out, proof_words,
                 row_tables);
    ZKM_API_END
}


// ---- column-layout handshake (include/zkm_b200.h): the offsets the kernels were compiled with
static std::vector<std::pair<uint32_t, uint32_t>> layout_pairs() {
    namespace c = tables::cpu;
    std::vector<std::pair<uint32_t, uint32_t>> v;
    for (int t = 0; t < tables::NUM_TABLE_KINDS; t++) v.push_back({(uint32_t)(ZKM_LK_NUM_COLUMNS + t), (uint32_t)tables::table_num_columns(t)});
    auto add = [&](zkm_layout_key_t k, int val) { v.push_back({(uint32_t)k, (uint32_t)val}); };
    add(ZKM_LK_CPU_IS_BOOTSTRAP_KERNEL, c::IS_BOOTSTRAP_KERNEL); add(ZKM_LK_CPU_IS_EXIT_KERNEL, c::IS_EXIT_KERNEL);
    add(ZKM_LK_CPU_CONTEXT, c::CONTEXT); add(ZKM_LK_CPU_CODE_CONTEXT, c::CODE_CONTEXT);
    add(ZKM_LK_CPU_PROGRAM_COUNTER, c::PROGRAM_COUNTER); add(ZKM_LK_CPU_NEXT_PROGRAM_COUNTER, c::NEXT_PROGRAM_COUNTER);
    add(ZKM_LK_CPU_IS_KERNEL_MODE, c::IS_KERNEL_MODE);
    add(ZKM_LK_CPU_OP_BINARY_OP, c::OP_BINARY_OP); add(ZKM_LK_CPU_OP_SYSCALL, c::OP_SYSCALL);
    add(ZKM_LK_CPU_BRANCH_SHOULD_JUMP, c::BR_SHOULD_JUMP); add(ZKM_LK_CPU_BRANCH_IS_NE, c::BR_IS_NE);
    add(ZKM_LK_CPU_OPCODE_BITS, c::OPCODE_BITS); add(ZKM_LK_CPU_RS_BITS, c::RS_BITS); add(ZKM_LK_CPU_RT_BITS, c::RT_BITS);
    add(ZKM_LK_CPU_RD_BITS, c::RD_BITS); add(ZKM_LK_CPU_SHAMT_BITS, c::SHAMT_BITS); add(ZKM_LK_CPU_FUNC_BITS, c::FUNC_BITS);
    add(ZKM_LK_CPU_IS_POSEIDON_SPONGE, c::IS_POSEIDON_SPONGE); add(ZKM_LK_CPU_IS_KECCAK_SPONGE, c::IS_KECCAK_SPONGE);
    add(ZKM_LK_CPU_IS_SHA_EXTEND_SPONGE, c::IS_SHA_EXTEND_SPONGE); add(ZKM_LK_CPU_IS_SHA_COMPRESS_SPONGE, c::IS_SHA_COMPRESS_SPONGE);
    add(ZKM_LK_CPU_GENERAL, c::GENERAL); add(ZKM_LK_CPU_MEMIO_IS_LH, c::MEMIO_IS_LH); add(ZKM_LK_CPU_MEMIO_AUX_FILTER, c::MEMIO_AUX_FILTER);
    add(ZKM_LK_CPU_CLOCK, c::CLOCK); add(ZKM_LK_CPU_MEM_CHANNELS, c::MEM_CHANNELS);
    add(ZKM_LK_CPU_MEM_CHANNEL_STRIDE, c::ch(1, 0) - c::ch(0, 0));
    add(ZKM_LK_CPU_CH_USED_REL, c::CH_USED); add(ZKM_LK_CPU_CH_IS_READ_REL, c::CH_IS_READ);
    add(ZKM_LK_CPU_CH_ADDR_CONTEXT_REL, c::CH_ADDR_CONTEXT); add(ZKM_LK_CPU_CH_ADDR_SEGMENT_REL, c::CH_ADDR_SEGMENT);
    add(ZKM_LK_CPU_CH_ADDR_VIRTUAL_REL, c::CH_ADDR_VIRTUAL); add(ZKM_LK_CPU_CH_VALUE_REL, c::CH_VALUE);
    const int G = c::GENERAL;
    add(ZKM_LK_CPU_G_SYSCALL_COND_REL, c::G_SYSCALL_COND - G); add(ZKM_LK_CPU_G_SYSCALL_SYSNUM_REL, c::G_SYSCALL_SYSNUM - G);
    add(ZKM_LK_CPU_G_SYSCALL_A0_REL, c::G_SYSCALL_A0 - G); add(ZKM_LK_CPU_G_SYSCALL_A1_REL, c::G_SYSCALL_A1 - G);
    add(ZKM_LK_CPU_G_MISC_RS_BITS_REL, c::G_MISC_RS_BITS - G); add(ZKM_LK_CPU_G_MISC_IS_MSB_REL, c::G_MISC_IS_MSB - G);
    add(ZKM_LK_CPU_G_MISC_IS_LSB_REL, c::G_MISC_IS_LSB - G); add(ZKM_LK_CPU_G_MISC_AUXM_REL, c::G_MISC_AUXM - G);
    add(ZKM_LK_CPU_G_MISC_AUXL_REL, c::G_MISC_AUXL - G); add(ZKM_LK_CPU_G_MISC_AUXS_REL, c::G_MISC_AUXS - G);
    add(ZKM_LK_CPU_G_MISC_RD_INDEX_REL, c::G_MISC_RD_INDEX - G); add(ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_0_REL, c::G_MISC_RD_INDEX_EQ_0 - G);
    add(ZKM_LK_CPU_G_MISC_RD_INDEX_EQ_29_REL, c::G_MISC_RD_INDEX_EQ_29 - G);
    add(ZKM_LK_CPU_G_IO_RS_LE_REL, c::G_IO_RS_LE - G); add(ZKM_LK_CPU_G_IO_RT_LE_REL, c::G_IO_RT_LE - G);
    add(ZKM_LK_CPU_G_IO_MEM_LE_REL, c::G_IO_MEM_LE - G); add(ZKM_LK_CPU_G_IO_AUX_RS0_MUL_RS1_REL, c::G_IO_AUX_RS0_MUL_RS1 - G);
    add(ZKM_LK_CPU_G_LOGIC_DIFF_PINV_REL, c::G_LOGIC_DIFF_PINV - G); add(ZKM_LK_CPU_G_HASH_VALUE_REL, c::G_HASH_VALUE - G);
    add(ZKM_LK_CPU_G_KHASH_VALUE_REL, c::G_KHASH_VALUE - G); add(ZKM_LK_CPU_G_SHASH_VALUE_REL, c::G_SHASH_VALUE - G);
    add(ZKM_LK_CPU_G_ELEMENT_VALUE_REL, c::G_ELEMENT_VALUE - G);
    return v;
}
int zkm_b200_layout_check(const uint32_t* pairs, size_t n_pairs, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(pairs || n_pairs == 0, "null argument");
    auto mine = layout_pairs();
    for (size_t i = 0; i < n_pairs; i++) {
        const uint32_t key = pairs[2 * i], val = pairs[2 * i + 1];
        bool found = false;
        for (auto& kv : mine)
            if (kv.first == key) {
                found = true;
                ZKM_CHECK(kv.second == val, "column layout mismatch at key " + std::to_string(key) + ": the caller has " + std::to_string(val) +
                                                ", the kernels were compiled with " + std::to_string(kv.second));
            }
        ZKM_CHECK(found, "unknown layout key " + std::to_string(key));
    }
    ZKM_API_END
}
int zkm_b200_layout_describe(uint32_t* pairs, size_t max_pairs, size_t* n_pairs, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(n_pairs, "null argument");
    auto mine = layout_pairs();
    *n_pairs = mine.size();
    for (size_t i = 0; i < mine.size() && i < max_pairs && pairs; i++) { pairs[2 * i] = mine[i].first; pairs[2 * i + 1] = mine[i].second; }
    ZKM_API_END
}

void zkm_b200_batch_free(zkm_batch_t* b) {
    if (!b) return;
    try { if (ctx_ready()) cudaStreamSynchronize(ctx().stream); } catch (...) {}
    delete b;
}

int zkm_b200_batch_get_coeffs(const zkm_batch_t* b, uint32_t col, uint64_t* out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(b && (int)col < b->b.ncols, "bad batch/column");
    b->b.coeffs.download(out, b->b.n(), (size_t)col * b->b.n());
    ZKM_API_END
}

int zkm_b200_batch_get_lde(const zkm_batch_t* b, uint32_t col, uint64_t* out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(b && (int)col < b->b.ncols, "bad batch/column");
    const Batch& B = b->b;
    size_t N = B.lde_n(), n = B.n();
    std::vector<u64> tmp(N);
    B.lde.download(tmp.data(), N, (size_t)col * N);
    int R = 1 << B.rate_bits;
    for (int j = 0; j < R; j++)
        for (size_t i = 0; i < n; i++) out[(i << B.rate_bits) | j] = tmp[(size_t)j * n + i];
    ZKM_API_END
}

int zkm_b200_batch_open(const zkm_batch_t* b, uint32_t leaf_index, uint64_t* leaf_out, uint64_t* siblings_out, char** err) {
    ZKM_API_BEGIN
    ZKM_CHECK(b, "null batch");
    const Batch& B = b->b;
    Ctx& c = ctx();
    ZKM_CHECK(leaf_index < B.lde_n(), "leaf index out of range");
    int path_len = B.tree.log_leaves - B.tree.cap_height;
    DevBuf idx(1, c.stream), rows(B.ncols, c.stream), path(path_len * 4 + 1, c.stream);
    u64 hidx = leaf_index;   // low 32 bits read as u32 on device (little endian)
    idx.upload(&hidx, 1);
    lde_gather_rows(B.lde.p, B.lde_n(), B.ncols, B.log_n, B.rate_bits, (const u32*)idx.p, 1, rows.p, c.stream);
    merkle_gather_paths(B.tree, (const u32*)idx.p, 1, path.p, c.stream);
    rows.download(leaf_out, B.ncols);
    if (path_len) path.download(siblings_out, (size_t)path_len * 4);
    ZKM_API_END
}

int zkm_b200_ntt(uint64_t* data, uint32_t ncols, uint32_t log_n, int kind, char** err) {
    ZKM_API_BEGIN
    Ctx& c = ctx();
    size_t n = (size_t)1 << log_n;
    DevBuf d((size_t)ncols * n, c.stream);
    d.upload(data, (size_t)ncols * n);
    if (kind == 0) ntt_forward(c.ntt, d.p, n, d.p, n, ncols, log_n, c.stream);
    else if (kind == 1) ntt_inverse(c.ntt, d.p, n, d.p, n, ncols, log_n, c.stream);
    else if (kind == 2) coset_intt(c.ntt, d.p, n, d.p, n, ncols, log_n, c.stream);
    else throw std::runtime_error("unknown transform kind");
    d.download(data, (size_t)ncols * n);
    ZKM_API_END
}

// SplitMix64 stream per column, reduced mod p (BASELINE.md §3 synthetic inputs): value i of column c =
// mix(seed_c + (i+1)*0x9E3779B97F4A7C15) mod p with seed_c = seed | c.
__global__ void synth_columns_kernel(u64* out, size_t n, u64 seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 c = blockIdx.y;
    u64 z = (seed | c) + (u64)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    out[c * n + i] = z >= GL_P ? z - GL_P : z;
}

int zkm_b200_synth_columns_device(uint64_t* d_out, uint32_t ncols, uint32_t log_n, uint64_t seed, char** err) {
    ZKM_API_BEGIN
    Ctx& c = ctx();
    size_t n = (size_t)1 << log_n;
    dim3 grid((unsigned)((n + 255) / 256), ncols);
    synth_columns_kernel<<<grid, 256, 0, c.stream>>>(d_out, n, seed);
    ZKM_LAUNCHED();
    ZKM_CUDA(stream_sync(c.stream));
    ZKM_API_END
}

// Synthetic segment traces: every cell uniform SplitMix64 mod p, except the columns read
// by CTL filters, which are drawn so that every Filter evaluates to 0 or 1 (at most one flag set per row,
// with probability 1/2), because get_helper_cols rejects non-binary filters (cross_table_lookup.rs:741).
// Such traces are not valid executions: the prover runs to completion, the proof does not verify.
__global__ void synth_flags_kernel(u64* out, size_t n, const int* flag_cols, int nflags, u64 seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 z = (seed ^ 0xF1A6F1A6F1A6F1A6ULL) + (u64)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    int hot = (z & 1) ? (int)((z >> 1) % (u64)nflags) : -1;
    for (int k = 0; k < nflags; k++) out[(size_t)flag_cols[k] * n + i] = (k == hot) ? 1 : 0;
}

static std::vector<int> filter_columns_of_table(const tables::System& sys, int t) {
    std::vector<int> cols;
    auto add_col = [&](const tables::Column& c) {
        for (auto& p : c.lin) cols.push_back(p.first);
        for (auto& p : c.next) cols.push_back(p.first);
    };
    auto add_filter = [&](const tables::Filter& f) {
        for (auto& pr : f.products) { add_col(pr.first); add_col(pr.second); }
        for (auto& c : f.constants) add_col(c);
    };
    for (auto& ctl : sys.ctls) {
        for (auto& lt : ctl.looking_tables) if (lt.table == t) add_filter(lt.filter);
        if (ctl.looked_table.table == t) add_filter(ctl.looked_table.filter);
    }
    for (auto& lk : tables::table_lookups(sys.kinds[t])) for (auto& f : lk.filter_columns) add_filter(f);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    return cols;
}

static void synth_trace_dev(int system_id, uint32_t table, uint32_t log_n, uint64_t seed, u64* d_out) {
    Ctx& c = ctx();
    tables::System sys = tables::make_system(system_id);
    ZKM_CHECK(table < sys.kinds.size(), "table index out of range");
    int ncols = tables::table_num_columns(sys.kinds[table]);
    size_t n = (size_t)1 << log_n;
    dim3 grid((unsigned)((n + 255) / 256), ncols);
    synth_columns_kernel<<<grid, 256, 0, c.stream>>>(d_out, n, seed);
    ZKM_LAUNCHED();
    std::vector<int> flags = filter_columns_of_table(sys, (int)table);
    if (!flags.empty()) {
        DevBuf df((flags.size() + 1) / 2 + 1, c.stream);
        ZKM_CUDA(cudaMemcpyAsync(df.p, flags.data(), flags.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
        synth_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(d_out, n, (const int*)df.p, (int)flags.size(), seed);
        ZKM_LAUNCHED();
        ZKM_CUDA(stream_sync(c.stream));
    }
    ZKM_CUDA(stream_sync(c.stream));
}

int zkm_b200_synth_trace_device(int system_id, uint32_t table, uint32_t log_n, uint64_t seed, uint64_t* d_out, char** err) {
    ZKM_API_BEGIN
    synth_trace_dev(system_id, table, log_n, seed, d_out);
    ZKM_API_END
}
int zkm_b200_synth_trace(int system_id, uint32_t table, uint32_t log_n, uint64_t seed, uint64_t* host_out, char** err) {
    ZKM_API_BEGIN
    tables::System sys = tables::make_system(system_id);
    ZKM_CHECK(table < sys.kinds.size(), "table index out of range");
    size_t total = (size_t)tables::table_num_columns(sys.kinds[table]) << log_n;
    DevBuf d(total, ctx().stream);
    synth_trace_dev(system_id, table, log_n, seed, d.p);
    d.download(host_out, total);
    ZKM_API_END
}
int zkm_b200_system_shape(int system_id, uint32_t* num_tables, uint32_t* ncols_out, uint32_t max_tables, char** err) {
    ZKM_API_BEGIN
    tables::System sys = tables::make_system(system_id);
    ZKM_CHECK(sys.kinds.size() <= max_tables, "ncols_out too small");
    *num_tables = (uint32_t)sys.kinds.size();
    for (size_t t = 0; t < sys.kinds.size(); t++) ncols_out[t] = (uint32_t)tables::table_num_columns(sys.kinds[t]);
    ZKM_API_END
}

__global__ void poseidon_states_kernel(u64* st, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = st[i * 12 + k];
    poseidon_permute_dev(s);
#pragma unroll
    for (int k = 0; k < 12; k++) st[i * 12 + k] = s[k];
}

int zkm_b200_poseidon_permute(uint64_t* states, size_t count, char** err) {
    ZKM_API_BEGIN
    Ctx& c = ctx();
    if (count) {
        DevBuf d(count * 12, c.stream);
        d.upload(states, count * 12);
        poseidon_states_kernel<<<(unsigned)((count + 127) / 128), 128, 0, c.stream>>>(d.p, count);
        ZKM_LAUNCHED();
        d.download(states, count * 12);
    }
    ZKM_API_END
}

int zkm_b200_transcript_permute(uint64_t* states, size_t count, char** err) {
    ZKM_API_BEGIN
    for (size_t i = 0; i < count; i++) poseidon_permute_host(states + 12 * i);
    ZKM_API_END
}

}  // extern "C"
