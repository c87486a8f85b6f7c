// int8 power-of-two twin of LightweightKWS (the esp-dl / esp_ppq export the firmware runs).
//
// Reference: ml_models/xiaoa.info (int8 weights :31-3136, per-tensor exponents :3139-3150, shipped known-answer
// vector :3153-3224), ml_models/xiaoa.json:5-20 (symmetric, per-tensor, power-of-two), device requantisation path
// main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:128-131,200-220.  Every tensor is int8 with a
// power-of-two scale; accumulation is int32; ReLU is applied before requantisation; requantisation is a right
// shift with round-half-to-even and int8 saturation; MaxPool works on int8; the global average is requantised
// from the exact mean.  Integer-exact: there is no floating-point tolerance in this path.
// Same thread mapping as cnn_fp32_kernel (lane <-> output channel, warp <-> 8-step time tile).
#pragma once
#include "ww_cnn.cuh"

namespace ww {

struct I8Weights {
    const signed char* w1t;  // [13][3][32]
    const signed char* w2t;  // [32][3][64]
    const signed char* w3t;  // [64][3][128]
    const signed char* fc1;  // [64][128]
    const signed char* fc2;  // [C][64]
    int num_classes;
    // right shifts (accumulator exponent - output exponent is negative: acc * 2^-shift)
    int sh1, sh2, sh3, shf1, shf2;
    int gap_num_shift;  // gap_q = rne(sum * 2^gap_num_shift / 7)
};

struct I8Args {
    const signed char* x;    // [n][13][63] int8 at the model-input exponent (coef-major)
    long long n_windows;
    signed char* out;        // [n][C] int8 at the output exponent
    I8Weights w;
};

// round-half-to-even right shift with int8 saturation
__device__ __forceinline__ int requant_i8(int acc, int shift) {
    int r;
    if (shift <= 0) {
        r = acc << (-shift);
    } else {
        const int half = 1 << (shift - 1);
        const int floor_q = acc >> shift;                 // arithmetic shift = floor
        const int rem = acc - (floor_q << shift);         // 0 .. 2^shift - 1
        r = floor_q + ((rem > half || (rem == half && (floor_q & 1))) ? 1 : 0);
    }
    return max(-128, min(127, r));
}

template <int CIN, int COUT, int IN_STRIDE>
__device__ __forceinline__ void conv8_i8(const signed char* __restrict__ wt, const int* __restrict__ xin, int o, int t0,
                                         int shift, int (&pooled)[4]) {
    int acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
#pragma unroll 4
    for (int c = 0; c < CIN; ++c) {
        const int w0 = wt[(c * 3 + 0) * COUT + o], w1 = wt[(c * 3 + 1) * COUT + o], w2 = wt[(c * 3 + 2) * COUT + o];
        const int* xr = xin + c * IN_STRIDE + t0;  // xr[i] = x[t0 + i - 1]
        int xs[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) xs[i] = xr[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += w0 * xs[i] + w1 * xs[i + 1] + w2 * xs[i + 2];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a0 = requant_i8(max(acc[2 * j], 0), shift), a1 = requant_i8(max(acc[2 * j + 1], 0), shift);
        pooled[j] = max(a0, a1);
    }
}

__global__ void __launch_bounds__(CNN_THREADS) cnn_i8_kernel(const I8Args a) {
    __shared__ int sm[CNN_SMEM_FLOATS];
    int* x0 = sm;                       // [13][68]  idx = t + 1
    int* a1 = x0 + 13 * X0_STRIDE;      // [32][36]
    int* a2 = a1 + 32 * A1_STRIDE;      // [64][20]
    int* gp = a2 + 64 * A2_STRIDE;      // [2][128]
    int* h1 = gp + 2 * 128;             // [64]
    const int tid = threadIdx.x, warp = warp_index_uniform(tid), lane = tid & 31;
    for (int i = tid; i < CNN_SMEM_FLOATS; i += CNN_THREADS) sm[i] = 0;
    __syncthreads();
    const int C = a.w.num_classes;
    for (long long win = blockIdx.x; win < a.n_windows; win += gridDim.x) {
        const signed char* src = a.x + win * (WW_N_MFCC * WW_WINDOW_FRAMES);
        for (int i = tid; i < WW_N_MFCC * WW_WINDOW_FRAMES; i += CNN_THREADS) {
            const int q = i / WW_WINDOW_FRAMES, t = i - q * WW_WINDOW_FRAMES;
            x0[q * X0_STRIDE + 1 + t] = src[i];
        }
        __syncthreads();
        int p[4];
        {
            const int o = lane, tg = warp;
            conv8_i8<13, 32, X0_STRIDE>(a.w.w1t, x0, o, 8 * tg, a.w.sh1, p);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tp = 4 * tg + j;
                a1[o * A1_STRIDE + 1 + tp] = tp < 31 ? p[j] : 0;
            }
        }
        __syncthreads();
        {
            const int o = lane + 32 * (warp & 1), tg = warp >> 1;
            conv8_i8<32, 64, A1_STRIDE>(a.w.w2t, a1, o, 8 * tg, a.w.sh2, p);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tp = 4 * tg + j;
                a2[o * A2_STRIDE + 1 + tp] = tp < 15 ? p[j] : 0;
            }
        }
        __syncthreads();
        {
            const int o = lane + 32 * (warp & 3), tg = warp >> 2;
            conv8_i8<64, 128, A2_STRIDE>(a.w.w3t, a2, o, 8 * tg, a.w.sh3, p);
            int s = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) s += (4 * tg + j < 7) ? p[j] : 0;
            gp[tg * 128 + o] = s;
        }
        __syncthreads();
        // global average: exact mean of 7 int8 values requantised (round half to even; ties cannot occur for /7)
        if (tid < 128) {
            const int s = (gp[tid] + gp[128 + tid]);
            const int num = a.w.gap_num_shift >= 0 ? (s << a.w.gap_num_shift) : s;  // numerator of the mean * 2^k
            const int den = a.w.gap_num_shift >= 0 ? 7 : (7 << (-a.w.gap_num_shift));
            // nearest integer to num/den (den odd multiple of 7 => no exact .5 unless den even; handle generally)
            int qv;
            {
                const long long n2 = 2LL * num, d2 = 2LL * den;
                long long fl = n2 >= 0 ? n2 / d2 : -((-n2 + d2 - 1) / d2);
                long long rem = n2 - fl * d2;  // 0 .. d2-1
                qv = (int)fl + ((rem > den || (rem == den && (fl & 1))) ? 1 : 0);
            }
            gp[tid] = max(-128, min(127, qv));
        }
        __syncthreads();
        {
            for (int oo = 0; oo < 8; ++oo) {
                const int o = warp * 8 + oo;
                int s = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) s += gp[lane + 32 * i] * (int)a.w.fc1[o * 128 + lane + 32 * i];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0) h1[o] = requant_i8(max(s, 0), a.w.shf1);
            }
        }
        __syncthreads();
        for (int c = warp; c < C; c += 8) {
            int s = h1[lane] * (int)a.w.fc2[c * 64 + lane] + h1[lane + 32] * (int)a.w.fc2[c * 64 + lane + 32];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) a.out[win * C + c] = (signed char)requant_i8(s, a.w.shf2);
        }
        __syncthreads();
    }
}

}  // namespace ww
