// Micro-benchmark + correctness check of the device Poseidon permutation variants (development tool).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I zkm_b200/csrc tools/micro/poseidon_bench.cu -o tools/micro/poseidon_bench
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#ifdef HAVE_V2
#include "poseidon_v2.cuh"
#endif
using namespace zkm;

template <int VARIANT>
__global__ void __launch_bounds__(128) k_perm(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) {
        if (VARIANT == 0) poseidon_permute(s);
#ifdef HAVE_V2
        else poseidon_permute_v2(s);
#endif
    }
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_perm_lb(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) poseidon_permute_v2(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}
// two states per thread: more independent work per warp
__global__ void __launch_bounds__(128) k_perm_x2(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= count) return;
    u64 s[12], t[12];
#pragma unroll
    for (int k = 0; k < 12; k++) { s[k] = in[k * count + i]; t[k] = in[k * count + i + 1]; }
    for (int r = 0; r < reps; r++) { poseidon_permute_v2(s); poseidon_permute_v2(t); }
#pragma unroll
    for (int k = 0; k < 12; k++) { out[k * count + i] = s[k]; out[k * count + i + 1] = t[k]; }
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_perm_v3(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) poseidon_permute_v3(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}
template <int V, int MINB>
__global__ void __launch_bounds__(128, MINB) k_perm_vx(const u64* __restrict__ in, u64* __restrict__ out, size_t count, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[k * count + i];
    for (int r = 0; r < reps; r++) {
        if (V == 4) poseidon_permute_v4(s);
        else if (V == 6) poseidon_permute_v6(s);
        else if (V == 8) poseidon_permute_v8(s);
        else if (V == 9) poseidon_permute_v9_t<true>(s);
        else if (V >= 20 && V < 28) poseidon_permute_v9_t<true, V - 20>(s);
        else poseidon_permute_v9_t<false>(s);
    }
#pragma unroll
    for (int k = 0; k < 12; k++) out[k * count + i] = s[k];
}
template <class F> void timeit(const char* name, F f, size_t count, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-28s %s %.3f ms -> %.3f Gperm/s\n", name, cudaGetErrorString(cudaGetLastError()), best, count * reps / (best * 1e-3) / 1e9);
}

int main() {
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
    std::vector<u64> ref(12 * 4096), got(12 * count);
    for (size_t i = 0; i < 4096; i++) {
        u64 s[12];
        for (int k = 0; k < 12; k++) s[k] = h[k * count + i];
        for (int r = 0; r < reps; r++) poseidon_permute(s);
        for (int k = 0; k < 12; k++) ref[k * 4096 + i] = s[k];
    }
    for (int variant = 0; variant < 2; variant++) {
#ifndef HAVE_V2
        if (variant == 1) break;
#endif
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0);
            if (variant == 0) k_perm<0><<<(unsigned)(count / 128), 128>>>(din, dout, count, reps);
            else k_perm<1><<<(unsigned)(count / 128), 128>>>(din, dout, count, reps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < 4096; i++) for (int k = 0; k < 12; k++) if (got[k * count + i] != ref[k * 4096 + i]) bad++;
        printf("variant %d: %s  %.3f ms for %zu perms -> %.3f Gperm/s   mismatches=%zu  first=%016llx\n", variant, cudaGetErrorString(e), best,
               count * reps, count * reps / (best * 1e-3) / 1e9, bad, (unsigned long long)got[0]);
    }
#ifdef HAVE_V2
    unsigned g = (unsigned)(count / 128);
    timeit("v2 lb(128,6)", [&] { k_perm_lb<6><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v2 lb(128,8)", [&] { k_perm_lb<8><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v2 lb(128,10)", [&] { k_perm_lb<10><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v2 lb(128,12)", [&] { k_perm_lb<12><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v3 fp64-mds lb(128,4)", [&] { k_perm_v3<4><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    {
        cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < 4096; i++) for (int k = 0; k < 12; k++) if (got[k * count + i] != ref[k * 4096 + i]) bad++;
        printf("v3 mismatches=%zu\n", bad);
    }
    timeit("v3 fp64-mds lb(128,6)", [&] { k_perm_v3<6><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v3 fp64-mds lb(128,8)", [&] { k_perm_v3<8><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    auto check = [&](const char* name) {
        cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < 4096; i++) for (int k = 0; k < 12; k++) if (got[k * count + i] != ref[k * 4096 + i]) bad++;
        printf("%s mismatches=%zu\n", name, bad);
    };
    timeit("v4 single-loop lb(128,6)", [&] { k_perm_vx<4, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v4");
    timeit("v4 single-loop lb(128,8)", [&] { k_perm_vx<4, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v6 rolled sbox lb(128,6)", [&] { k_perm_vx<6, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v6");
    timeit("v6 rolled sbox lb(128,8)", [&] { k_perm_vx<6, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v8 merged consts lb(128,6)", [&] { k_perm_vx<8, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v8");
    timeit("v9 freq-mds paired lb(128,6)", [&] { k_perm_vx<9, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9");
    timeit("v9 freq-mds paired lb(128,5)", [&] { k_perm_vx<9, 5><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v9 freq-mds paired lb(128,4)", [&] { k_perm_vx<9, 4><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v9 freq-mds paired lb(128,8)", [&] { k_perm_vx<9, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v9 cv1 (I2F split) lb(128,8)", [&] { k_perm_vx<21, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9cv1");
    timeit("v9 cv2 (F2I recombine) lb(128,8)", [&] { k_perm_vx<22, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9cv2");
    timeit("v9 cv3 (I2F+F2I) lb(128,8)", [&] { k_perm_vx<23, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9cv3");
    timeit("v9 cv4 (mad.wide pack) lb(128,8)", [&] { k_perm_vx<24, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9cv4");
    timeit("v9 cv6 (mad.wide+F2I) lb(128,8)", [&] { k_perm_vx<26, 8><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v9cv6");
    timeit("v9 cv3 (I2F+F2I) lb(128,6)", [&] { k_perm_vx<23, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v10 freq-mds unpaired lb(128,6)", [&] { k_perm_vx<10, 6><<<g, 128>>>(din, dout, count, reps); }, count, reps); check("v10");
    timeit("v10 freq-mds unpaired lb(128,5)", [&] { k_perm_vx<10, 5><<<g, 128>>>(din, dout, count, reps); }, count, reps);
    timeit("v2 x2 states/thread", [&] { k_perm_x2<<<g / 2, 128>>>(din, dout, count, reps); }, count, reps);
#endif
    return 0;
}
