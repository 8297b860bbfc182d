// Micro-benchmark + correctness check of the device Poseidon permutation (development tool; A/B knobs of poseidon_v2.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I zkm_b200/csrc [-DZKM_P9_SBOX_UNROLL=1] \
//             tools/micro/poseidon_bench.cu -o tools/micro/poseidon_bench
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon_v2.cuh"
using namespace zkm;

template <int MINB, int CV>
__global__ void __launch_bounds__(128, MINB) k_v9(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) poseidon_permute_v9_t<true, CV>(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_v9_threads(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) poseidon_permute_v9_t<true, 3>(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}
// two states per thread, advanced in lockstep (poseidon_permute_v9_x2)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_v9_x2(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= count) return;
    const size_t half = count / 2;
    u64 a[12], b[12];
#pragma unroll
    for (int k = 0; k < 12; k++) { a[k] = in[k * count + i]; b[k] = in[k * count + half + i]; }
    for (int r = 0; r < reps; r++) poseidon_permute_v9_x2<3>(a, b);
#pragma unroll
    for (int k = 0; k < 12; k++) { out[k * count + i] = a[k]; out[k * count + half + i] = b[k]; }
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;                  // only the product configuration (A/B builds of one knob)
    const size_t count = 1 << 20;
    const int reps = 8;
    std::vector<u64> h(12 * count);
    u64 z = 12345;
    for (auto& x : h) { z = z * 6364136223846793005ULL + 1442695040888963407ULL; u64 v = z ^ (z >> 29); x = v >= GL_P ? v - GL_P : v; }
    for (int k = 0; k < 12; k++) h[k * count + 0] = 0;                 // KAT rows
    for (int k = 0; k < 12; k++) h[k * count + 1] = k;
    for (int k = 0; k < 12; k++) h[k * count + 2] = GL_P - 1;
    u64 *din, *dout;
    cudaMalloc(&din, h.size() * 8); cudaMalloc(&dout, h.size() * 8);
    cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const size_t NREF = 4096;
    std::vector<u64> ref(12 * 2 * NREF), got(12 * count);
    auto ref_of = [&](size_t i, size_t slot) {
        u64 s[12];
        for (int k = 0; k < 12; k++) s[k] = h[k * count + i];
        for (int r = 0; r < reps; r++) poseidon_permute(s);          // portable host/device reference (poseidon.cuh)
        for (int k = 0; k < 12; k++) ref[k * 2 * NREF + slot] = s[k];
    };
    for (size_t i = 0; i < NREF; i++) { ref_of(i, i); ref_of(count / 2 + i, NREF + i); }
    auto run = [&](const char* name, auto launch) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemset(dout, 0, got.size() * 8);
        float best = 1e9;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < NREF; i++) for (int k = 0; k < 12; k++) {
            if (got[k * count + i] != ref[k * 2 * NREF + i]) bad++;
            if (got[k * count + count / 2 + i] != ref[k * 2 * NREF + NREF + i]) bad++;
        }
        printf("%-40s %s %8.3f ms -> %.3f Gperm/s  mismatches=%zu\n", name, cudaGetErrorString(e), best, count * reps / (best * 1e-3) / 1e9, bad);
    };
    const unsigned g = (unsigned)(count / 128);
    printf("ZKM_P9_SBOX_UNROLL=%d ZKM_P9_MULMIX=%d\n", ZKM_P9_SBOX_UNROLL, ZKM_P9_MULMIX);
    run("v9 cv3 lb(128,8)  [product]", [&] { k_v9<8, 3><<<g, 128>>>(din, dout, count, reps); });
    if (quick) { run("v9 cv3 lb(128,6)", [&] { k_v9<6, 3><<<g, 128>>>(din, dout, count, reps); }); return 0; }
    run("v9 cv3 lb(128,6)", [&] { k_v9<6, 3><<<g, 128>>>(din, dout, count, reps); });
    run("v9 cv3 lb(128,4)", [&] { k_v9<4, 3><<<g, 128>>>(din, dout, count, reps); });
    run("v9 cv0 lb(128,8)", [&] { k_v9<8, 0><<<g, 128>>>(din, dout, count, reps); });
    run("v9 cv1 lb(128,8)", [&] { k_v9<8, 1><<<g, 128>>>(din, dout, count, reps); });
    run("v9 cv2 lb(128,8)", [&] { k_v9<8, 2><<<g, 128>>>(din, dout, count, reps); });
    run("v9 cv3 256 threads lb(256,4)", [&] { k_v9_threads<256, 4><<<(unsigned)(count / 256), 256>>>(din, dout, count, reps); });
    run("v9 cv3 64 threads lb(64,16)", [&] { k_v9_threads<64, 16><<<(unsigned)(count / 64), 64>>>(din, dout, count, reps); });
    run("v9 x2 states/thread lb(128,4)", [&] { k_v9_x2<4><<<g / 2, 128>>>(din, dout, count, reps); });
    run("v9 x2 states/thread lb(128,3)", [&] { k_v9_x2<3><<<g / 2, 128>>>(din, dout, count, reps); });
    run("v9 x2 states/thread lb(128,2)", [&] { k_v9_x2<2><<<g / 2, 128>>>(din, dout, count, reps); });
    return 0;
}
