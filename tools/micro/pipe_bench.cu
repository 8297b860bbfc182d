// Throughput probes for the integer instructions the Goldilocks kernels are made of (development tool).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) probe(u64* out, u32 a, u32 c) {
    u64 acc[8]; u32 x[8], y[8], w[8]; double dd[8]; double dc = 1.0000001 + c * 1e-9, da = a * 1e-3;
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = threadIdx.x + k; x[k] = threadIdx.x * 7 + k; y[k] = threadIdx.x * 3 + k; w[k] = k; dd[k] = threadIdx.x + k; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));          // IMAD.WIDE acc
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c), "r"(a));                  // IMAD 32
            if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c));                                 // IADD3
            if (MODE == 3) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a)); }                              // 1 IMAD.WIDE : 1 IADD
            if (MODE == 4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
                             asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[k]) : "r"(c)); }                              // 1 : 2
            if (MODE == 5) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[k]), "r"(c)); acc[k] ^= t; } // IMAD.WIDE RZ + 2 LOP3
            if (MODE == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }                              // independent add
            if (MODE == 8) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a));
                             asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[k]) : "r"(y[k])); }                           // 1 wide : 2 alu independent
            if (MODE == 9) { asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(y[k]), "+r"(w[k]) : "r"(a)); }   // carry chain pair
            if (MODE == 10) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[k]), "r"(y[k]));
                             asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(y[k]), "+r"(w[k]) : "r"((u32)t), "r"((u32)(t >> 32))); } // mul.wide + 64-bit add
            if (MODE == 11) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[k]) : "r"(y[k])); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(w[k])); } // alu only
            if (MODE == 12) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[k]) : "d"(dc), "d"(da));                      // DFMA
            if (MODE == 13) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[k]) : "d"(dc), "d"(da));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }                             // DFMA + IADD
            if (MODE == 14) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[k]) : "d"(dc), "d"(da));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c)); }       // DFMA + IMAD.WIDE
            if (MODE == 15) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[k]) : "d"(dc), "d"(da));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }                             // DFMA + IMAD.WIDE + IADD
            if (MODE == 16) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(y[k]) : "r"(w[k]), "r"(a));           // SHF
            if (MODE == 17) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[k]) : "r"(w[k]), "r"(a));          // LOP3
            if (MODE == 18) asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(y[k]) : "r"(a));                           // IMAD.IADD form
            if (MODE == 19) { asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(y[k]) : "r"(a));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(w[k]) : "r"(c)); }                             // IMAD.IADD + IADD3 (1:1)
            if (MODE == 20) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(c));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(y[k]), "r"(c));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(w[k]) : "r"(a)); }                             // 2 wide : 1 alu
            if (MODE == 21) { asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(w[k]) : "r"(c));
                              asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(c)); }                             // 3 indep IADD
            if (MODE == 22) { asm volatile("sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 %3, 0, 0;" : "+r"(y[k]), "+r"(w[k]) : "r"(a), "r"(x[k])); } // 3-long borrow chain
            if (MODE == 6) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c), "r"(a));
                             u32 y = (u32)acc[k]; asm volatile("add.u32 %0, %0, %1;" : "+r"(y) : "r"(a)); acc[k] = y; } // 1 IMAD32 : 1 IADD
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k] + x[k] + y[k] + w[k] + (u64)dd[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_iter) {
    u64* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, 3, 5);
    cudaEventRecord(e0); probe<MODE><<<148 * 8, 256>>>(d, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = 148.0 * 8 * 256 * ITERS * 8 * per_iter;
    printf("%-34s %.3f ms  %.2f T thread-instr/s  (%.1f per SM per clk at 1.965 GHz)\n", name, ms, instr / (ms * 1e-3) / 1e12,
           instr / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    run<0>("IMAD.WIDE accumulate", 1);
    run<1>("IMAD 32-bit", 1);
    run<2>("IADD3", 1);
    run<3>("IMAD.WIDE acc + IADD (1:1)", 2);
    run<4>("IMAD.WIDE acc + IADD + LOP (1:2)", 3);
    run<5>("IMAD.WIDE(RZ) + 2 LOP3", 3);
    run<6>("IMAD32 + IADD (1:1)", 2);
    run<7>("IMAD.WIDE acc + indep IADD (1:1)", 2);
    run<8>("IMAD.WIDE acc + 2 indep ALU (1:2)", 3);
    run<9>("add.cc/addc pair", 2);
    run<10>("mul.wide + add.cc/addc (1:2)", 3);
    run<11>("xor + add (ALU only)", 2);
    run<12>("DFMA", 1);
    run<13>("DFMA + IADD (1:1)", 2);
    run<14>("DFMA + IMAD.WIDE (1:1)", 2);
    run<15>("DFMA + IMAD.WIDE + IADD (1:1:1)", 3);
    run<16>("SHF", 1);
    run<17>("LOP3", 1);
    run<18>("IMAD x*1+y", 1);
    run<19>("IMAD x*1+y + IADD3 (1:1)", 2);
    run<20>("2 IMAD.WIDE + 1 IADD", 3);
    run<21>("3 indep IADD", 3);
    run<22>("3-long borrow chain", 3);
    return 0;
}
