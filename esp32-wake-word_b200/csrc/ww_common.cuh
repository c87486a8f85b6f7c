// Shared device helpers for the wake-word hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WW_N_FFT 512
#define WW_WIN 320
#define WW_HOP 256
#define WW_N_BINS 257
#define WW_N_MELS 40
#define WW_N_MFCC 13
#define WW_WINDOW_FRAMES 63

namespace ww {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// ---- packed complex arithmetic (sm_100 f32x2 family: SASS FADD2 / FMUL2 / FFMA2) -------------------------
// A complex value lives in one aligned 64-bit register pair {re (lo), im (hi)}.  One packed instruction does
// both halves and takes one issue slot (measured on B200, tools/ubench/f32x2_rate.cu: 2 packed warp-instructions
// per clock per SM = the same 128 results/clk/SM as 4 scalar ones), and ptxas folds half swaps, per-half
// negations and scalar broadcasts of the operands into the instruction's .LO_HI / .NP / .F32 modifiers, so a
// complex add is 1 instruction, a complex multiply 2, and "t -+ i d" 1 -- half the scalar count.
struct cpx {
    unsigned long long v;
};
__device__ __forceinline__ cpx cpk(float re, float im) {
    cpx r;
    asm("mov.b64 %0,{%1,%2};" : "=l"(r.v) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ float2 cunpk(cpx a) {
    float2 r;
    asm("mov.b64 {%0,%1},%2;" : "=f"(r.x), "=f"(r.y) : "l"(a.v));
    return r;
}
__device__ __forceinline__ float cre(cpx a) { return cunpk(a).x; }
__device__ __forceinline__ float cim(cpx a) { return cunpk(a).y; }
__device__ __forceinline__ cpx cswap(cpx a) {
    const float2 f = cunpk(a);
    return cpk(f.y, f.x);
}
__device__ __forceinline__ cpx p_neg(cpx a) {
    const float2 f = cunpk(a);
    return cpk(-f.x, -f.y);
}
__device__ __forceinline__ cpx p_conj(cpx a) {
    const float2 f = cunpk(a);
    return cpk(f.x, -f.y);
}
__device__ __forceinline__ cpx p_add(cpx a, cpx b) {
    cpx r;
    asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ cpx p_sub(cpx a, cpx b) {
    cpx r;
    asm("sub.rn.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ cpx p_mul(cpx a, cpx b) {
    cpx r;
    asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ cpx p_fma(cpx a, cpx b, cpx c) {
    cpx r;
    asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// t + (-i) d   and   t - (-i) d
__device__ __forceinline__ cpx p_add_mi(cpx t, cpx d) { return p_fma(cswap(d), cpk(1.f, -1.f), t); }
__device__ __forceinline__ cpx p_sub_mi(cpx t, cpx d) { return p_fma(cswap(d), cpk(-1.f, 1.f), t); }
// a * (-i)
__device__ __forceinline__ cpx p_mul_mi(cpx a) { return p_mul(cswap(a), cpk(1.f, -1.f)); }
// a * (c + i s)
__device__ __forceinline__ cpx p_cmul(cpx a, float c, float s);
__device__ __forceinline__ cpx p_cmulc(cpx a, cpx w) {
    const float2 f = cunpk(w);
    return p_cmul(a, f.x, f.y);
}
__device__ __forceinline__ cpx p_cmul(cpx a, float c, float s) { return p_fma(cswap(a), cpk(-s, s), p_mul(a, cpk(c, c))); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk copy (TMA) -----------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WW_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WW_DONE_%=;\n"
        "bra WW_WAIT_%=;\n"
        "WW_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// the same copy with an L2 cache policy (createpolicy): evict-first for data that is read exactly once
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// asynchronous prefetch of a global range into L2 (16-byte aligned address, size a multiple of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// Programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization): a kernel lets the next launch of
// its stream start early with launch_dependents; the next kernel may run its prologue (tables, weights, TMEM, its first
// input block) and calls pdl_wait before it touches anything the previous kernel writes or still reads -- the wait
// returns when that kernel has completed and its memory is visible.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Warp index of a thread, taken out of a shuffle: to the compiler threadIdx.x >> 5 is a per-thread value, so every loop
// bound, branch and address derived from it counts as divergent -- each __shfl_sync / __ballot_sync inside such a loop is
// then compiled with a WARPSYNC + ENDCOLLECTIVE slow path, and operands that must live in uniform registers (UMMA
// descriptors) are moved there one instruction at a time.  A value that comes out of a shuffle with a constant source
// lane is known to be warp-uniform.
__device__ __forceinline__ int warp_index_uniform(int tid) { return __shfl_sync(0xffffffffu, tid >> 5, 0); }
__device__ __forceinline__ int warp_index_uniform() { return warp_index_uniform((int)threadIdx.x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ww
