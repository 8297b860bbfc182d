// Row-major trace rows -> column-major trace columns on the device.
// Replaces reference util.rs:37-47 trace_rows_to_poly_values (the CPU transposition every row-generated table goes through
// in Traces::into_tables, witness/traces.rs:274-305): the host hands over the rows exactly as its generators leave them
// (Vec<[F; COLUMNS]>, one contiguous n x ncols block) and never builds the column vectors.
// 32 x 32 tiles through shared memory: both the row-major reads and the column-major writes are 256-byte coalesced runs.
#include "dev.cuh"

namespace zkm {

__global__ void transpose_rows_kernel(const u64* __restrict__ rows, u64* __restrict__ cols, size_t n, int ncols) {
    __shared__ u64 tile[32][33];
    const size_t r0 = (size_t)blockIdx.x * 32;          // row tiles on grid.x (up to 2^31 - 1), column tiles on grid.y (<= 65535)
    const int c0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int c = c0 + threadIdx.x;
        if (c < ncols) tile[k][threadIdx.x] = rows[(r0 + k) * (size_t)ncols + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int c = c0 + k;
        if (c < ncols) cols[(size_t)c * n + r0 + threadIdx.x] = tile[threadIdx.x][k];
    }
}

// n is a power of two >= 32 (trace heights are >= 64: all_stark.rs:115)
void transpose_rows_to_cols(const u64* rows, u64* cols, size_t n, int ncols, cudaStream_t s) {
    ZKM_CHECK(n >= 32 && (n & (n - 1)) == 0, "transpose: height must be a power of two >= 32");
    ProfScope ps("transpose_rows", s, 16.0 * (double)n * ncols);
    ZKM_CHECK(n / 32 <= 0x7fffffffu && (ncols + 31) / 32 <= 65535, "transpose: table too large");
    dim3 grid((unsigned)(n / 32), (unsigned)((ncols + 31) / 32));
    transpose_rows_kernel<<<grid, dim3(32, 8), 0, s>>>(rows, cols, n, ncols);
    ZKM_LAUNCHED();
}

}  // namespace zkm

// ---- Arithmetic table: range-check columns on the device.
// Replaces reference arithmetic_stark.rs:127-153 generate_range_checks: RANGE_COUNTER = 0, 1, ..., 2^16 - 1 and then constant,
// RC_FREQUENCIES[x] += number of cells of the 18 shared columns equal to x.  `bad` is raised if a shared cell is >= 2^16
// (the reference asserts "column value ... exceeds the max range value").
namespace zkm {

__global__ void arith_range_counter_kernel(u64* cols, size_t n, int counter_col, u64 range_max) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cols[(size_t)counter_col * n + i] = i < range_max ? (u64)i : range_max - 1;
}
__global__ void arith_frequencies_kernel(u64* cols, size_t n, int first_shared, int num_shared, int freq_col, u64 range_max,
                                         unsigned* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (size_t)num_shared) return;
    u64 x = cols[(size_t)first_shared * n + i];             // the shared columns are contiguous: column-major block
    if (x >= range_max) { atomicExch(bad, 1u); return; }
    atomicAdd((unsigned long long*)&cols[(size_t)freq_col * n + x], 1ULL);
}

void arith_generate_range_checks(u64* cols, size_t n, int first_shared, int num_shared, int counter_col, int freq_col,
                                 unsigned* d_bad, cudaStream_t s) {
    const u64 range_max = 1ull << 16;
    ZKM_CHECK(n >= range_max, "arithmetic table shorter than the range check (2^16 rows)");
    ProfScope ps("arith_range_checks", s, 8.0 * (double)n * (num_shared + 2));
    ZKM_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned), s));
    arith_range_counter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cols, n, counter_col, range_max);
    ZKM_LAUNCHED();
    size_t cells = n * (size_t)num_shared;
    arith_frequencies_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(cols, n, first_shared, num_shared, freq_col, range_max, d_bad);
    ZKM_LAUNCHED();
}

}  // namespace zkm
