// Issue-rate probe for the packed fp32 instructions of sm_100a (FADD2 / FFMA2) against their scalar forms.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE> __global__ void probe(float *out, int iters, float seed)
{
    float s[8]; u64 p[8]; unsigned n[8];
    for (int i = 0; i < 8; ++i) { s[i] = seed + i; p[i] = ((u64)__float_as_uint(seed + i) << 32) | __float_as_uint(seed - i); n[i] = i; }
    const u64 k = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) s[i] = add1(s[i], seed);             // 8 FADD
            if (MODE == 1) p[i] = add2(p[i], k);                // 8 FADD2
            if (MODE == 2) s[i] = fma1(s[i], seed, seed);       // 8 FFMA
            if (MODE == 3) p[i] = fma2(p[i], k, k);             // 8 FFMA2
            if (MODE == 4) { s[i] = add1(s[i], seed); n[i] = iadd(n[i], 3u); }  // 8 FADD + 8 IADD
            if (MODE == 5) { p[i] = add2(p[i], k); n[i] = iadd(n[i], 3u); }     // 8 FADD2 + 8 IADD
            if (MODE == 6) n[i] = iadd(n[i], 3u);               // 8 IADD
            if (MODE == 7) { p[i] = add2(p[i], k); s[i] = add1(s[i], seed); }   // 8 FADD2 + 8 FADD
        }
    }
    float acc = 0; for (int i = 0; i < 8; ++i) acc += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + n[i];
    if (acc == 12345.678f) out[0] = acc;
}
template <int MODE> void run(const char *name, int instr_per_iter)
{
    float *out; cudaMalloc(&out, 4);
    const int iters = 20000, blocks = 148 * 8, threads = 256;
    probe<MODE><<<blocks, threads>>>(out, 100, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); probe<MODE><<<blocks, threads>>>(out, iters, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)blocks * (threads / 32) * iters * instr_per_iter;
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-22s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, warp_instr / cycles / 148.0, clk_khz / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("FADD", 8); run<1>("FADD2", 8); run<2>("FFMA", 8); run<3>("FFMA2", 8);
    run<4>("FADD+IADD", 16); run<5>("FADD2+IADD", 16); run<6>("IADD", 8); run<7>("FADD2+FADD", 16);
    return 0;
}
