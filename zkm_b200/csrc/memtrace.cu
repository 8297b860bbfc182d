// Memory table generation on the device.
// Replaces reference memory/memory_stark.rs:133-244 MemoryStark::generate_trace: sort the memory operations by
// (context, segment, virt, timestamp) (stable), fill_gaps (:186-217: dummy reads so that every range-checked delta stays
// below the table height), pad to a power of two with copies of the last operation (:219-237), then the first-change
// flags, the range-check column (:83-131), COUNTER and FREQUENCIES (:152-162); writes to register 0 are recorded as 0
// (:62-72).  Output: the 13 columns, column-major, ready for the prover.
//
// Layout of one operation on input (7 words): context, segment, virt, timestamp, is_read, value, filter.
// Sort: bitonic network on (address key, timestamp | index) pairs -- the index makes every key unique, which is what a
// stable sort needs from an unstable network.  O(n log^2 n) compare-exchanges on 16-byte keys: ~6 ms for 2^20 operations.
#include "dev.cuh"

namespace zkm {

namespace {
constexpr int MEM_COLS = 13;
enum { C_FILTER = 0, C_TS, C_IS_READ, C_CTX, C_SEG, C_VIRT, C_VALUE, C_CTX_FC, C_SEG_FC, C_VIRT_FC, C_RC, C_COUNTER, C_FREQ };
constexpr u64 SEG_REGISTER_FILE = 4;

struct Key { u64 hi, lo; };
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

__global__ void mem_keys_kernel(const u64* ops, size_t n_ops, size_t n1, Key* keys) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    Key k = {~0ull, ~0ull};
    if (i < n_ops) {
        const u64* o = ops + 7 * i;
        k.hi = (o[0] << 40) | (o[1] << 32) | o[2];
        k.lo = (o[3] << 24) | (u64)i;
    }
    keys[i] = k;
}
__global__ void bitonic_step_kernel(Key* keys, size_t n1, size_t j, size_t k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    size_t l = i ^ j;
    if (l <= i) return;
    Key a = keys[i], b = keys[l];
    bool up = (i & k) == 0;
    if (key_less(b, a) == up) { keys[i] = b; keys[l] = a; }
}
// number of dummy operations fill_gaps inserts after sorted operation i
__global__ void mem_gaps_kernel(const u64* ops, const Key* keys, size_t n_ops, u64 max_rc, u64* gaps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ops) return;
    u64 g = 0;
    if (i + 1 < n_ops) {
        const u64* a = ops + 7 * (keys[i].lo & 0xFFFFFF);
        const u64* b = ops + 7 * (keys[i + 1].lo & 0xFFFFFF);
        if (a[0] == b[0] && a[1] == b[1]) {
            if (a[2] != b[2]) g = (b[2] - a[2] - 1) / (max_rc + 1);
            else { u64 d = b[3] - a[3]; g = d > max_rc ? (d - 1) / max_rc : 0; }
        }
    }
    gaps[i] = g;
}
// exclusive scan of `gaps` by one CTA (n <= 2^24: a few thousand 1024-wide chunks); out[n] = total, *last_gap = last i with gaps > 0
__global__ void mem_scan_kernel(const u64* gaps, size_t n, u64* out, long long* last_gap) {
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry;
    __shared__ long long last;
    if (threadIdx.x == 0) { carry = 0; last = -1; }
    __syncthreads();
    for (size_t base = 0; base < n; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        u64 v = i < n ? gaps[i] : 0;
        u64 x = v;
        for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            u64 w = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        u64 before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < n) { out[i] = before; if (v) atomicMax(&last, (long long)i); }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[n] = carry; *last_gap = last; }
}
struct MemRow { u64 filter, ts, is_read, ctx, seg, virt, value; };
__device__ __forceinline__ void put_row(u64* cols, size_t n2, size_t r, const MemRow& m) {
    u64 value = m.value;
    if (!m.is_read && m.ctx == 0 && m.seg == SEG_REGISTER_FILE && m.virt == 0) value = 0;      // writes to R0 are recorded as 0
    cols[C_FILTER * n2 + r] = m.filter; cols[C_TS * n2 + r] = m.ts; cols[C_IS_READ * n2 + r] = m.is_read;
    cols[C_CTX * n2 + r] = m.ctx; cols[C_SEG * n2 + r] = m.seg; cols[C_VIRT * n2 + r] = m.virt; cols[C_VALUE * n2 + r] = value;
}
// One thread per OUTPUT row: row r belongs to the sorted operation i with the largest start position pos(i) <= r (binary
// search over the monotone positions pos(i) = i + scan[i] + (i > last_gap ? pad : 0)); offset 0 is the operation itself,
// offsets 1..gaps[i] are its fill_gaps dummy reads, and the rows after those (only behind the last element of the pre-padding
// list) are padding copies.  Every row is written by its own thread, coalesced per column, so neither a long padding tail
// (up to n2/2 rows) nor a huge address gap serialises in one thread.
__device__ __forceinline__ size_t mem_row_pos(const u64* scan, size_t i, long long last_gap, size_t pad) {
    return i + scan[i] + ((last_gap >= 0 && (long long)i > last_gap) ? pad : 0);
}
__global__ void mem_place_kernel(const u64* ops, const Key* keys, const u64* gaps, const u64* scan, size_t n_ops, u64 max_rc,
                                 long long last_gap, size_t pad, size_t n2, u64* cols) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n2) return;
    size_t lo = 0, hi = n_ops - 1;                      // pos(0) = 0 <= r
    while (lo < hi) {
        size_t mid = (lo + hi + 1) >> 1;
        if (mem_row_pos(scan, mid, last_gap, pad) <= r) lo = mid; else hi = mid - 1;
    }
    const size_t i = lo;
    const u64 j = r - mem_row_pos(scan, i, last_gap, pad);
    const u64* o = ops + 7 * (keys[i].lo & 0xFFFFFF);
    MemRow m = {o[6], o[3], o[4], o[0], o[1], o[2], o[5]};
    if (j == 0) { put_row(cols, n2, r, m); return; }
    const u64 g = gaps[i];
    MemRow d = m;
    d.filter = 0; d.is_read = 1;
    // dummy number jj of this operation (padding = copies of the last element of the list before padding: the last dummy
    // pushed if there is one, else the last sorted operation as a filtered-off read)
    const u64 jj = j <= g ? j : g;
    if (jj) {
        const u64* nx = ops + 7 * (keys[i + 1].lo & 0xFFFFFF);
        if (o[2] != nx[2]) { d.virt = m.virt + jj * (max_rc + 1); d.ts = 0; d.value = 0; }
        else d.ts = m.ts + jj * max_rc;
    }
    put_row(cols, n2, r, d);
}
__global__ void mem_flags_kernel(u64* cols, size_t n2, unsigned* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    cols[C_COUNTER * n2 + i] = i;
    u64 cfc = 0, sfc = 0, vfc = 0, rc = 0;
    if (i + 1 < n2) {
        u64 c0 = cols[C_CTX * n2 + i], c1 = cols[C_CTX * n2 + i + 1], s0 = cols[C_SEG * n2 + i], s1 = cols[C_SEG * n2 + i + 1];
        u64 v0 = cols[C_VIRT * n2 + i], v1 = cols[C_VIRT * n2 + i + 1], t0 = cols[C_TS * n2 + i], t1 = cols[C_TS * n2 + i + 1];
        cfc = c0 != c1; sfc = (s0 != s1) && !cfc; vfc = (v0 != v1) && !sfc && !cfc;
        rc = cfc ? c1 - c0 - 1 : sfc ? s1 - s0 - 1 : vfc ? v1 - v0 - 1 : t1 - t0;
        if (rc >= n2) { atomicExch(bad, 1u); rc = 0; }
    }
    cols[C_CTX_FC * n2 + i] = cfc; cols[C_SEG_FC * n2 + i] = sfc; cols[C_VIRT_FC * n2 + i] = vfc; cols[C_RC * n2 + i] = rc;
    atomicAdd((unsigned long long*)&cols[C_FREQ * n2 + rc], 1ULL);
}
}  // namespace

// ops: n_ops x 7 words on the host.  Returns the table height; `cols` receives 13 x height words (column-major, device).
size_t memory_generate_trace_dev(const u64* h_ops, size_t n_ops, DevBuf& cols, cudaStream_t s) {
    ZKM_CHECK(n_ops >= 1, "No memory ops?");
    ZKM_CHECK(n_ops < ((size_t)1 << 24), "too many memory operations");
    for (size_t i = 0; i < n_ops; i++) {
        const u64* o = h_ops + 7 * i;
        // value: the reference stores it with from_canonical_u32 (memory_stark.rs:62-72)
        ZKM_CHECK(o[0] < (1ull << 24) && o[1] < 256 && o[2] < (1ull << 32) && o[3] < (1ull << 40) && o[4] <= 1 && o[5] < (1ull << 32) &&
                  o[6] <= 1, "memory operation out of range");
    }
    size_t n1 = 1;
    while (n1 < n_ops) n1 <<= 1;
    const u64 max_rc = n1 - 1;                              // memory_ops.len().next_power_of_two() - 1
    DevBuf ops(7 * n_ops, s), keys(2 * n1, s), gaps(n_ops, s), scan(n_ops + 1, s), misc(2, s);
    ops.upload(h_ops, 7 * n_ops);
    Key* k = (Key*)keys.p;
    const unsigned th = 256;
    ProfScope ps("memory_trace", s, 8.0 * 7 * (double)n_ops);
    mem_keys_kernel<<<(unsigned)((n1 + th - 1) / th), th, 0, s>>>(ops.p, n_ops, n1, k);
    ZKM_LAUNCHED();
    for (size_t kk = 2; kk <= n1; kk <<= 1)
        for (size_t j = kk >> 1; j > 0; j >>= 1) {
            bitonic_step_kernel<<<(unsigned)((n1 + th - 1) / th), th, 0, s>>>(k, n1, j, kk);
            ZKM_LAUNCHED();
        }
    mem_gaps_kernel<<<(unsigned)((n_ops + th - 1) / th), th, 0, s>>>(ops.p, k, n_ops, max_rc, gaps.p);
    ZKM_LAUNCHED();
    mem_scan_kernel<<<1, 1024, 0, s>>>(gaps.p, n_ops, scan.p, (long long*)misc.p);
    ZKM_LAUNCHED();
    u64 total = 0;
    long long last_gap = -1;
    scan.download(&total, 1, n_ops);
    misc.download((u64*)&last_gap, 1, 0);
    const size_t n_list = n_ops + (size_t)total;
    size_t n2 = 1;
    while (n2 < n_list) n2 <<= 1;
    cols.alloc((size_t)MEM_COLS * n2, s);
    cols.zero();
    mem_place_kernel<<<(unsigned)((n2 + th - 1) / th), th, 0, s>>>(ops.p, k, gaps.p, scan.p, n_ops, max_rc, last_gap, n2 - n_list, n2, cols.p);
    ZKM_LAUNCHED();
    unsigned* d_bad = (unsigned*)(misc.p + 1);
    ZKM_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), s));
    mem_flags_kernel<<<(unsigned)((n2 + th - 1) / th), th, 0, s>>>(cols.p, n2, d_bad);
    ZKM_LAUNCHED();
    u64 bad = 0;
    misc.download(&bad, 1, 1);
    ZKM_CHECK((unsigned)bad == 0, "Range check is too large. Bug in fill_gaps?");
    return n2;
}

}  // namespace zkm
