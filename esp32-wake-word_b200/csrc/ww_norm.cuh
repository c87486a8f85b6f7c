// normalize_mfcc for rows of ANY length (ml_models/src/extract_mfcc.py:47-88), sm_100a.
//
// The 63-frame windows of the hot path are normalised inside the CNN kernels (or by cmvn_rows_kernel); this kernel
// serves the reference's general call shape -- normalize_mfcc(mfcc[n_mfcc, T], method) over the time axis for any T:
//   'standardization' / 'cmvn'  (x - mean) / (std + 1e-8), torch.std = unbiased (N - 1), std == 0 -> 1   (:61-78)
//   'minmax'                    (x - min) / (max - min + 1e-8)                                              (:66-70)
// One group of G threads (a warp for short rows, a 256-thread CTA for long ones) per row; the row is read twice
// (statistics, then the normalised write), the second time out of L1/L2.  Algorithmic bytes 8 T per row.
#pragma once
#include "ww_common.cuh"

namespace ww {

enum { NORM_STANDARD = 0, NORM_MINMAX = 1 };

struct NormArgs {
    const float* x;        // [n_rows][row_stride], T valid values per row
    float* out;            // same geometry
    long long n_rows;
    long long row_stride;
    int T;
    int method;
};

template <int G>
__device__ __forceinline__ float group_sum(float v, float* red) {
    v = warp_sum(v);
    if constexpr (G > 32) {
        const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
        __syncthreads();  // red is reused between reductions
        if (lane == 0) red[warp] = v;
        __syncthreads();
        v = 0.f;
#pragma unroll
        for (int i = 0; i < G / 32; ++i) v += red[i];  // same order in every thread
    }
    return v;
}
template <int G, bool MAX>
__device__ __forceinline__ float group_ext(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = MAX ? fmaxf(v, w) : fminf(v, w);
    }
    if constexpr (G > 32) {
        const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        v = red[0];
#pragma unroll
        for (int i = 1; i < G / 32; ++i) v = MAX ? fmaxf(v, red[i]) : fminf(v, red[i]);
    }
    return v;
}

// G = 32: 8 rows per 256-thread CTA (a warp each); G = 256: one row per CTA
template <int G>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const NormArgs a) {
    __shared__ float red[8];
    const int rows_per_cta = 256 / G;
    const int sub = G == 32 ? warp_index_uniform() : 0;
    const int li = G == 32 ? (threadIdx.x & 31) : threadIdx.x;
    // G = 256 uses __syncthreads in the reductions: every thread of the CTA walks the same rows
    for (long long row = (long long)blockIdx.x * rows_per_cta + sub; row < a.n_rows; row += (long long)gridDim.x * rows_per_cta) {
        const float* x = a.x + row * a.row_stride;
        float* o = a.out + row * a.row_stride;
        const int T = a.T;
        if (a.method == NORM_MINMAX) {
            float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
            for (int t = li; t < T; t += G) {
                const float v = x[t];
                lo = fminf(lo, v);
                hi = fmaxf(hi, v);
            }
            lo = group_ext<G, false>(lo, red);
            hi = group_ext<G, true>(hi, red);
            const float den = (hi - lo) + 1e-8f;
            for (int t = li; t < T; t += G) o[t] = (x[t] - lo) / den;
        } else {
            float s = 0.f;
            for (int t = li; t < T; t += G) s += x[t];
            const float mean = group_sum<G>(s, red) / (float)T;
            float ss = 0.f;
            for (int t = li; t < T; t += G) {
                const float d = x[t] - mean;
                ss = fmaf(d, d, ss);
            }
            ss = group_sum<G>(ss, red);
            // torch.std of a single value is NaN and NaN == 0 is false: the reference then divides by NaN
            float sd = T > 1 ? sqrtf(ss / (float)(T - 1)) : __int_as_float(0x7fc00000);
            if (sd == 0.f) sd = 1.f;
            const float den = sd + 1e-8f;
            for (int t = li; t < T; t += G) o[t] = (x[t] - mean) / den;
        }
    }
}

}  // namespace ww
