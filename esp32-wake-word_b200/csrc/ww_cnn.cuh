// CMVN + LightweightKWS forward + decision, fp32 on the CUDA cores (exact-arithmetic path).
//
// Replaces  normalize_mfcc(mfcc,'cmvn')                 ml_models/src/extract_mfcc.py:47-88
//           detect_task CMVN (device twin)              esp_wake_word_detector.cpp:179-211
//           LightweightKWS.forward                      ml_models/src/wakeModel.py:29-34
//           torch.sigmoid(out) > 0.5 / sigmoid*100>=80  ml_models/main.py:53, esp_wake_word_detector.cpp:226-245
//
// One CTA (256 threads) scores one 63-frame window at a time and loops over windows (persistent grid).
// Windows are addressed through three strides, so a batch of clips [B,13,63] and the sliding windows of a
// stream's [13,T] feature plane (window stride = 1 frame) use the same kernel.
// Layer mapping: lane <-> output channel, warp <-> (channel block, 8-step time tile); each thread keeps
// 8 accumulators, activations are warp-broadcast from shared memory with 128-bit loads, weights are read
// coalesced (layout [cin][tap][cout]) through L1.  ReLU + MaxPool(2) are applied in registers.
// This kernel is the bit-stable fp32 reference path of the engine: the tensor-core path re-scores its
// borderline windows with it.
#pragma once
#include "ww_common.cuh"

namespace ww {

enum { CMVN_NONE = 0, CMVN_PY = 1, CMVN_DEVICE = 2 };
enum { DECIDE_NONE = 0, DECIDE_LOGIT = 1, DECIDE_DEVICE = 2 };

struct CnnWeights {
    const float* w1t;  // [13][3][32]
    const float* w2t;  // [32][3][64]
    const float* w3t;  // [64][3][128]
    const float* fc1;  // [64][128]
    const float* fc2;  // [C][64]
    int num_classes;
};

struct CnnArgs {
    const float* feats;      // feats[win*win_stride + coef*coef_stride + frame*frame_stride]
    long long win_stride;
    long long coef_stride;
    long long frame_stride;
    long long n_windows;
    long long group_windows;   // 0: flat addressing; else window id = group*group_windows + j and the window
    long long group_stride;    //    lives at feats + group*group_stride + j*win_stride (concurrent streams)
    const long long* index;  // optional: window ids to score (re-score list); nullptr = 0..n_windows-1
    const int* index_count;  // optional device count for `index` (n_windows is then the capacity)
    unsigned long long* index_total;  // optional running total of re-scored windows (block 0 adds index_count)
    int index_compact;       // 1: the features of index[k] are window k of `feats` (the fused clip kernel's compact copy)
    int cmvn_mode;
    int decide_mode;
    float threshold;         // DECIDE_LOGIT: logit > threshold; DECIDE_DEVICE: sigmoid*100 >= threshold
    float* logits;           // [n_windows][C]
    unsigned char* decisions;  // [n_windows] (class 0), may be null
    float* norm_out;         // optional [n_windows][13][63] normalised features (ww_cmvn), may be null
    CnnWeights w;
};

constexpr int CNN_THREADS = 256;
constexpr int X0_STRIDE = 68, A1_STRIDE = 36, A2_STRIDE = 20;
constexpr int CNN_SMEM_FLOATS = 13 * X0_STRIDE + 32 * A1_STRIDE + 64 * A2_STRIDE + 2 * 128 + 64;

// lroundf (half away from zero) + saturation to int8.  rintf is one instruction (half to even); only an exact tie needs
// the other neighbour, and v - rintf(v) is exact, so the test is exact too (checked against round-half-away on 5 M values,
// every half-integer in +-300, +-inf and the 0.49999997 case: identical)
__device__ __forceinline__ float lround_clamp_i8(float v) {
    float r = rintf(v);
    if (fabsf(v - r) == 0.5f) r = v + copysignf(0.5f, v);
    return fminf(fmaxf(r, -128.f), 127.f);
}

// The device-style CMVN (esp_wake_word_detector.cpp:179-211) divides: mean = sum / 63 and (v - mean) / (std + 1e-8),
// each followed by a rounding to int8.  Both kernels must reproduce the IEEE quotients bit for bit (the int8 step turns
// a last-bit difference into a whole unit), but they need not execute a division to do so:
//  * x / 63: r = x c with c = RN(1/63), corrected once with the exact residual fma(-r, 63, x), is the correctly rounded
//    quotient (Markstein's division step: c is the correctly rounded reciprocal and 63's significand is not all ones;
//    checked for every integer sum in [-8064, 8064] and 2e8 random significands over 53 binades: no mismatch);
//  * lround(d / den): q = d * rcp(den) (MUFU, 1 ulp) is within 2^-22 |q| of the correctly rounded quotient, so for
//    |q| <= 256 the two round to the same integer unless q lies within 1e-4 of k + 1/2 -- only then (2 in 10 000) the
//    division is done.
__device__ __forceinline__ float div63_exact(float x) {
    const float c = 1.f / 63.f;
    const float r = x * c;
    return fmaf(fmaf(-r, 63.f, x), c, r);
}
__device__ __forceinline__ float rcp_approx(float x) {   // 1 ulp, one MUFU
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float div_lround_clamp_i8(float d, float den, float rinv) {
    const float q = d * rinv;
    const float r = rintf(q);
    float out = fminf(fmaxf(r, -128.f), 127.f);
    // negated comparisons: a NaN quotient takes the exact path as well
    if (!(fabsf(q) > 256.f) && !(fabsf(fabsf(q - r) - 0.5f) > 1.0e-4f)) out = lround_clamp_i8(d / den);
    return out;
}

// conv(k=3, pad=1, no bias) + ReLU + MaxPool(2) for 8 consecutive output steps of one channel.
template <int CIN, int COUT, int IN_STRIDE>
__device__ __forceinline__ void conv8(const float* __restrict__ wt, const float* __restrict__ xin, int o, int t0,
                                      float (&pooled)[4]) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int c = 0; c < CIN; ++c) {
        const float w0 = __ldg(wt + (c * 3 + 0) * COUT + o);
        const float w1 = __ldg(wt + (c * 3 + 1) * COUT + o);
        const float w2 = __ldg(wt + (c * 3 + 2) * COUT + o);
        const float* xr = xin + c * IN_STRIDE + t0;  // xr[i] = x[t0 + i - 1]
        const float4 a = *reinterpret_cast<const float4*>(xr);
        const float4 b = *reinterpret_cast<const float4*>(xr + 4);
        const float2 d = *reinterpret_cast<const float2*>(xr + 8);
        const float xs[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = fmaf(w0, xs[i], acc[i]);
            acc[i] = fmaf(w1, xs[i + 1], acc[i]);
            acc[i] = fmaf(w2, xs[i + 2], acc[i]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) pooled[j] = fmaxf(fmaxf(acc[2 * j], acc[2 * j + 1]), 0.f);
}

__global__ void __launch_bounds__(CNN_THREADS) cnn_fp32_kernel(const __grid_constant__ CnnArgs a) {
    __shared__ __align__(16) float sm[CNN_SMEM_FLOATS];
    float* x0 = sm;                        // [13][68]  idx = t + 1
    float* a1 = x0 + 13 * X0_STRIDE;       // [32][36]
    float* a2 = a1 + 32 * A1_STRIDE;       // [64][20]
    float* gp = a2 + 64 * A2_STRIDE;       // [2][128]
    float* h1 = gp + 2 * 128;              // [64]

    const int tid = threadIdx.x, warp = warp_index_uniform(tid), lane = tid & 31;
    for (int i = tid; i < CNN_SMEM_FLOATS; i += CNN_THREADS) sm[i] = 0.f;
    __syncthreads();

    long long n = a.n_windows;
    if (a.index_count) {
        const long long c = *a.index_count;
        n = c < n ? c : n;
        if (a.index_total && blockIdx.x == 0 && tid == 0) atomicAdd(a.index_total, (unsigned long long)n);
    }
    const int C = a.w.num_classes;
    n = __shfl_sync(0xffffffffu, n, 0);   // a loaded loop bound: tell the compiler it is warp-uniform (see warp_index_uniform)

    for (long long it = blockIdx.x; it < n; it += gridDim.x) {
        const long long win = a.index ? a.index[it] : it;
        const float* src = a.index_compact ? a.feats + it * a.win_stride
                           : a.group_windows
                               ? a.feats + (win / a.group_windows) * a.group_stride + (win % a.group_windows) * a.win_stride
                               : a.feats + win * a.win_stride;

        // ---- CMVN: warp handles coefficients warp, warp+8 ----
        for (int q = warp; q < WW_N_MFCC; q += 8) {
            const float* row = src + q * a.coef_stride;
            float v0 = row[lane * a.frame_stride];
            float v1 = (lane + 32 < WW_WINDOW_FRAMES) ? row[(lane + 32) * a.frame_stride] : 0.f;
            const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
            if (a.cmvn_mode == CMVN_DEVICE) {
                v0 = lround_clamp_i8(v0);
                v1 = has1 ? lround_clamp_i8(v1) : 0.f;
            }
            float z0 = v0, z1 = v1;
            if (a.cmvn_mode == CMVN_PY) {
                // Operation for operation the arithmetic of the tensor-core kernel's CMVN (tc_cmvn_py: same reduction
                // tree, mean and 1/(std + eps) as multiplications by correctly rounded reciprocals), so that both kernels
                // feed the network the SAME z bit for bit.  That matters for rows that are constant up to rounding
                // (digital silence: c0 = -87.377 in every frame): there (x - mean) / (std + eps) is rounding noise over
                // rounding noise, a last-bit difference in the mean changes z by O(1), and only identical arithmetic
                // keeps the tensor path inside its guard band of this kernel.  Within 1 ulp of the reference's
                // (x - mean) / (std + 1e-8) everywhere else (normalize_mfcc itself is cmvn_rows_kernel: true divisions).
                const float mean = warp_sum(v0 + v1) * (1.f / (float)WW_WINDOW_FRAMES);
                const float d0 = v0 - mean, d1 = has1 ? v1 - mean : 0.f;
                const float ss = warp_sum(fmaf(d0, d0, d1 * d1));
                float sd = sqrtf(ss * (1.f / (float)(WW_WINDOW_FRAMES - 1)));
                if (sd == 0.f) sd = 1.f;
                const float inv = __frcp_rn(sd + 1e-8f);
                z0 = d0 * inv;
                z1 = d1 * inv;
            } else if (a.cmvn_mode != CMVN_NONE) {
                const float mean = div63_exact(warp_sum(v0 + v1));   // the sum of 63 int8 values is exact in any order
                const float d0 = v0 - mean, d1 = has1 ? v1 - mean : 0.f;
                const float ss = warp_sum(d0 * d0 + d1 * d1);
                {
                    const float den = sqrtf(div63_exact(ss)) + 1e-8f;
                    const float rinv = rcp_approx(den);
                    // int8 at exponent 0 -> model input at exponent -4: saturates at 127/16
                    z0 = fminf(fmaxf(div_lround_clamp_i8(d0, den, rinv) * 16.f, -128.f), 127.f) * 0.0625f;
                    z1 = fminf(fmaxf(div_lround_clamp_i8(d1, den, rinv) * 16.f, -128.f), 127.f) * 0.0625f;
                }
            }
            x0[q * X0_STRIDE + 1 + lane] = z0;
            if (has1) x0[q * X0_STRIDE + 33 + lane] = z1;
            if (a.norm_out) {
                float* no = a.norm_out + win * (WW_N_MFCC * WW_WINDOW_FRAMES) + q * WW_WINDOW_FRAMES;
                no[lane] = z0;
                if (has1) no[lane + 32] = z1;
            }
        }
        __syncthreads();
        if (a.logits == nullptr) continue;  // CMVN-only call

        float p[4];
        // ---- conv1: 13 -> 32, T 63 -> 31 ----
        {
            const int o = lane, tg = warp;
            conv8<13, 32, X0_STRIDE>(a.w.w1t, x0, o, 8 * tg, p);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tp = 4 * tg + j;
                a1[o * A1_STRIDE + 1 + tp] = tp < 31 ? p[j] : 0.f;
            }
        }
        __syncthreads();
        // ---- conv2: 32 -> 64, T 31 -> 15 ----
        {
            const int o = lane + 32 * (warp & 1), tg = warp >> 1;
            conv8<32, 64, A1_STRIDE>(a.w.w2t, a1, o, 8 * tg, p);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tp = 4 * tg + j;
                a2[o * A2_STRIDE + 1 + tp] = tp < 15 ? p[j] : 0.f;
            }
        }
        __syncthreads();
        // ---- conv3: 64 -> 128, T 15 -> 7, then global average pool ----
        {
            const int o = lane + 32 * (warp & 3), tg = warp >> 2;
            conv8<64, 128, A2_STRIDE>(a.w.w3t, a2, o, 8 * tg, p);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) s += (4 * tg + j < 7) ? p[j] : 0.f;
            gp[tg * 128 + o] = s;
        }
        __syncthreads();
        // ---- fc1: 128 -> 64 + ReLU (warp handles 8 outputs, lanes over k) ----
        {
            float g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = lane + 32 * i;
                g[i] = (gp[k] + gp[128 + k]) / 7.f;
            }
#pragma unroll
            for (int oo = 0; oo < 8; ++oo) {
                const int o = warp * 8 + oo;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) s = fmaf(g[i], __ldg(a.w.fc1 + o * 128 + lane + 32 * i), s);
                s = warp_sum(s);
                if (lane == 0) h1[o] = fmaxf(s, 0.f);
            }
        }
        __syncthreads();
        // ---- fc2: 64 -> C, decision on class 0 ----
        for (int c = warp; c < C; c += 8) {
            float s = h1[lane] * __ldg(a.w.fc2 + c * 64 + lane);
            s = fmaf(h1[lane + 32], __ldg(a.w.fc2 + c * 64 + lane + 32), s);
            s = warp_sum(s);
            if (lane == 0) {
                a.logits[win * C + c] = s;
                if (c == 0 && a.decisions) {
                    unsigned char d = 0;
                    if (a.decide_mode == DECIDE_LOGIT) d = s > a.threshold;
                    else if (a.decide_mode == DECIDE_DEVICE) d = (1.f / (1.f + expf(-s)) * 100.f) >= a.threshold;
                    a.decisions[win] = d;
                }
            }
        }
        // x0/a1/a2 are rewritten only after the next window's barriers; h1/gp reads are complete before
        // the next window's conv3/fc1 writes because three __syncthreads() separate them.
    }
}

}  // namespace ww
