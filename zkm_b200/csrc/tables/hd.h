// Host/device qualifiers and constant-table plumbing shared by the constraint templates.
// The templates in tables/*.h are written ONCE, generic over the value type P (the analogue of the
// reference's `eval_packed_generic<FE, P, D2>`, prover/src/stark.rs:41-47): the CUDA quotient kernel
// instantiates them with the device Goldilocks type, the CPU oracle with its own scalar and
// quadratic-extension types.  P must provide: P(u64 canonical constant), + - * (binary), unary -.
#pragma once
#include <stdint.h>

#ifndef ZKM_HD
#ifdef __CUDACC__
#define ZKM_HD __host__ __device__ __forceinline__
#else
#define ZKM_HD inline
#endif
#endif

// ZKM_DEF_CONST(name, N, init): a read-only u64 table visible from host and device code as ZKM_K(name).
#ifdef __CUDACC__
#define ZKM_DEF_CONST(name, N, ...)                      \
    static const uint64_t H_##name[N] = __VA_ARGS__;     \
    static __device__ __constant__ const uint64_t D_##name[N] = __VA_ARGS__;
#else
#define ZKM_DEF_CONST(name, N, ...) static const uint64_t H_##name[N] = __VA_ARGS__;
#endif
#ifdef __CUDA_ARCH__
#define ZKM_K(name) D_##name
#else
#define ZKM_K(name) H_##name
#endif

// Loops over bits / limbs / channels in the constraint templates are kept rolled on the device: the quotient
// kernels are straight-line code executed once per thread, so their cost is instruction fetch, i.e. code
// size (the fully unrolled Cpu kernel was 87 000 SASS instructions).
#ifdef __CUDACC__
#define ZKM_ROLLED _Pragma("unroll 1")
#else
#define ZKM_ROLLED
#endif
