// Microbenchmark: issue rate of scalar FP32 (FADD/FMUL/FFMA) against the packed sm_100 forms
// (add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// Prints lane-results per clock per SM (a packed instruction produces 64 results per warp).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

#define ILP 8
template <int MODE> __global__ void __launch_bounds__(256) kern(float* out, int iters, float seed) {
    float s[ILP]; u64 p[ILP];
    float c0 = seed, c1 = seed * 0.5f;
    u64 pc0 = pk(c0, c1), pc1 = pk(c1, c0);
#pragma unroll
    for (int i = 0; i < ILP; i++) { s[i] = threadIdx.x * 0.001f + i; p[i] = pk(s[i], s[i] + 1.f); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) s[i] = fadd(s[i], c0);
                if (MODE == 1) s[i] = fmul(s[i], c0);
                if (MODE == 2) s[i] = ffma(s[i], c0, c1);
                if (MODE == 3) p[i] = add2(p[i], pc0);
                if (MODE == 4) p[i] = mul2(p[i], pc0);
                if (MODE == 5) p[i] = fma2(p[i], pc0, pc1);
                if (MODE == 6) p[i] = sub2(p[i], pc0);
                if (MODE == 7) { if (i & 1) s[i] = fadd(s[i], c0); else p[i] = add2(p[i], pc0); }      // mix scalar+packed
                if (MODE == 8) s[i] = ffma(s[i], 1.0009765625f, c1);                                   // immediate form
                if (MODE == 9) { if (i & 1) s[i] = fadd(s[i], c0); else s[i] = ffma(s[i], c0, c1); }  // FADD+FFMA mix
                if (MODE == 10) { p[i] = add2(p[i], p[(i + 1) % ILP]); }                               // reg-reg packed
                if (MODE == 11) { s[i] = fadd(s[i], s[(i + 1) % ILP]); }                               // reg-reg scalar
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc += s[i]; float lo, hi; asm("mov.b64 {%0,%1},%2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); acc += lo + hi; }
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE> void run(const char* name, int results_per_inst, float* d, int sms, int warps_per_sm) {
    int iters = 4096; int threads = 256; int blocks = sms * (warps_per_sm / 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<blocks, threads>>>(d, 64, 1.0001f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); kern<MODE><<<blocks, threads>>>(d, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double inst_per_warp = (double)iters * 4 * ILP;
    double warp_inst = inst_per_warp * blocks * threads / 32;
    double cycles = best * 1e-3 * clk_khz * 1e3;
    printf("%-28s warps/SM %2d  %.3f ms  warp-inst/clk/SM %.3f  results/clk/SM %.1f (at %d kHz nominal)\n", name, warps_per_sm, best,
           warp_inst / cycles / sms, warp_inst * results_per_inst / cycles / sms, clk_khz);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 4096);
    for (int w : {8, 16, 32}) {
        run<0>("FADD reg,const", 32, d, sms, w);
        run<11>("FADD reg,reg", 32, d, sms, w);
        run<1>("FMUL", 32, d, sms, w);
        run<2>("FFMA 3-reg", 32, d, sms, w);
        run<8>("FFMA imm", 32, d, sms, w);
        run<9>("FADD+FFMA mix", 32, d, sms, w);
        run<3>("FADD2", 64, d, sms, w);
        run<10>("FADD2 reg,reg", 64, d, sms, w);
        run<6>("FSUB2 (sub.f32x2)", 64, d, sms, w);
        run<4>("FMUL2", 64, d, sms, w);
        run<5>("FFMA2", 64, d, sms, w);
        run<7>("FADD + FADD2 mix", 48, d, sms, w);
    }
    return 0;
}
