// Row-major trace rows -> column-major trace columns on the device.
// Replaces reference util.rs:37-47 trace_rows_to_poly_values (the CPU transposition every row-generated table goes through
// in Traces::into_tables, witness/traces.rs:274-305): the host hands over the rows exactly as its generators leave them
// (Vec<[F; COLUMNS]>, one contiguous n x ncols block) and never builds the column vectors.
// 32 x 32 tiles through shared memory: both the row-major reads and the column-major writes are 256-byte coalesced runs.
#include "dev.cuh"

namespace zkm {

__global__ void transpose_rows_kernel(const u64* __restrict__ rows, u64* __restrict__ cols, size_t n, int ncols) {
    __shared__ u64 tile[32][33];
    const size_t r0 = (size_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
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
    dim3 grid((unsigned)((ncols + 31) / 32), (unsigned)(n / 32));
    transpose_rows_kernel<<<grid, dim3(32, 8), 0, s>>>(rows, cols, n, ncols);
    ZKM_LAUNCHED();
}

}  // namespace zkm
