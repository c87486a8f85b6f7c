// Front-of-frontend DSP, sm_100a.
//
//   tdm_downmix_kernel  4-channel TDM 48 kHz int16 -> mono 16 kHz int16, the integer arithmetic of record_task
//                       (main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:103-121): bit-exact
//   augment_kernel      augment_audio_waveform (ml_models/src/extract_mfcc.py:90-121): the five deterministic
//                       variants of a padded clip (original, speed 0.8 / 1.2 by linear interpolation, volume 0.7 / 1.3)
//
// Both are streaming byte/element kernels bound by HBM: 26 B (24 in + 2 out) per 16 kHz sample for the down-mix,
// 4 B in + 20 B out per sample for the augmentation.  Loads and stores are 16-byte (8-byte) vectors, fully
// coalesced and marked streaming (.cs); the grids are a multiple of the SM count and walk the work persistently.
#pragma once
#include "ww_common.cuh"

namespace ww {

struct TdmArgs {
    const int16_t* tdm;      // [n_signals][in_stride] : frames of 4 interleaved int16 channels at 48 kHz
    long long in_stride;     // int16 elements between signals
    int16_t* out;            // [n_signals][out_stride] mono 16 kHz
    long long out_stride;
    long long n_signals;
    long long n_out;         // 16 kHz samples per signal (each consumes 3 TDM frames = 12 int16)
    int vec_ok;              // 1: 16-byte loads / 8-byte stores are aligned for every signal
};

// one TDM frame {CH0 MIC-L, CH1 AEC ref, CH2 MIC-R, CH3 unused} -> mono, cpp:103-111
//   weighted = (L << 6) + (ref << 5) + (R << 6);  mono = (int16_t)(weighted >> 7)   (the cast wraps)
__device__ __forceinline__ int tdm_mix(int l, int ref, int r) {
    const int weighted = l * 64 + ref * 32 + r * 64;
    return (int)(int16_t)(weighted >> 7);
}
// [1, 2, 1] / 4 decimator over three consecutive mono samples, cpp:114-121
__device__ __forceinline__ int16_t tdm_decim(int m0, int m1, int m2) { return (int16_t)((m0 + 2 * m1 + m2) >> 2); }

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)(int16_t)(w >> 16); }

constexpr int TDM_THREADS = 256;
constexpr int TDM_TILE_OUT = TDM_THREADS * 4;   // 1024 output samples = 3072 TDM frames = 24 576 B per tile

// Vector path: a CTA walks tiles of 1024 output samples.  The 24 KB of TDM frames of a tile are read with fully
// coalesced 16-byte streaming loads (thread t takes chunks t, t+256, ... : 512 B contiguous per warp instruction),
// mixed to mono right away (2 frames per chunk), and the mono samples are parked in 6 KB of shared memory; after
// one barrier thread t decimates outputs 4t..4t+3 from its 12 consecutive mono samples (three conflict-free LDS.64)
// and stores them as one 8-byte word (256 B contiguous per warp).  A thread that read its own 96 contiguous bytes
// instead would fetch 16 B per L1 wavefront and cap the kernel at ~60 % of HBM bandwidth (measured).
__global__ void __launch_bounds__(TDM_THREADS) tdm_downmix_kernel(const TdmArgs a) {
    __shared__ __align__(16) uint32_t mono_s[2][TDM_THREADS * 6];   // 12 mono int16 per thread, double buffered
    const int tid = threadIdx.x;
    if (a.vec_ok) {
        const long long n_vec = a.n_out & ~3LL;                     // outputs covered by the vector path
        const long long tiles_per_sig = (n_vec + TDM_TILE_OUT - 1) / TDM_TILE_OUT;
        const long long total = a.n_signals * tiles_per_sig;
        int buf = 0;
        for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
            const long long sig = tile / tiles_per_sig, tl = tile - sig * tiles_per_sig;
            const long long out0 = tl * TDM_TILE_OUT;
            const long long n_here = n_vec - out0 < TDM_TILE_OUT ? n_vec - out0 : TDM_TILE_OUT;  // multiple of 4
            const int chunks = (int)(n_here / 4) * 6;               // 16-byte chunks (2 frames each) in this tile
            const uint4* src = reinterpret_cast<const uint4*>(a.tdm + sig * a.in_stride + out0 * 12);
            uint4 v[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int c = i * TDM_THREADS + tid;
                v[i] = c < chunks ? __ldcs(src + c) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                // uint4 = two frames {L|ref, R|ch3, L|ref, R|ch3}
                const uint32_t m0 = (uint16_t)tdm_mix(lo16(v[i].x), hi16(v[i].x), lo16(v[i].y));
                const uint32_t m1 = (uint16_t)tdm_mix(lo16(v[i].z), hi16(v[i].z), lo16(v[i].w));
                mono_s[buf][i * TDM_THREADS + tid] = m0 | (m1 << 16);
            }
            __syncthreads();   // the other buffer is reused only after the NEXT barrier: one barrier per tile
            if (tid * 4 < n_here) {
                const uint2* mp = reinterpret_cast<const uint2*>(&mono_s[buf][tid * 6]);
                const uint2 p0 = mp[0], p1 = mp[1], p2 = mp[2];
                const int m[12] = {lo16(p0.x), hi16(p0.x), lo16(p0.y), hi16(p0.y), lo16(p1.x), hi16(p1.x),
                                   lo16(p1.y), hi16(p1.y), lo16(p2.x), hi16(p2.x), lo16(p2.y), hi16(p2.y)};
                const uint32_t o0 = (uint16_t)tdm_decim(m[0], m[1], m[2]), o1 = (uint16_t)tdm_decim(m[3], m[4], m[5]);
                const uint32_t o2 = (uint16_t)tdm_decim(m[6], m[7], m[8]), o3 = (uint16_t)tdm_decim(m[9], m[10], m[11]);
                __stcs(reinterpret_cast<uint2*>(a.out + sig * a.out_stride + out0) + tid, make_uint2(o0 | (o1 << 16), o2 | (o3 << 16)));
            }
        }
        // scalar tail: the last n_out % 4 samples of every signal
        const int tail = (int)(a.n_out - n_vec);
        for (long long it = (long long)blockIdx.x * blockDim.x + tid; it < a.n_signals * tail; it += (long long)gridDim.x * blockDim.x) {
            const long long sig = it / tail, k = n_vec + (it - sig * tail);
            const int16_t* fr = a.tdm + sig * a.in_stride + k * 12;
            a.out[sig * a.out_stride + k] =
                tdm_decim(tdm_mix(fr[0], fr[1], fr[2]), tdm_mix(fr[4], fr[5], fr[6]), tdm_mix(fr[8], fr[9], fr[10]));
        }
        return;
    }
    // unaligned buffers: one output sample per thread, scalar loads
    const long long total = a.n_signals * a.n_out;
    for (long long it = (long long)blockIdx.x * blockDim.x + tid; it < total; it += (long long)gridDim.x * blockDim.x) {
        const long long sig = it / a.n_out, k = it - sig * a.n_out;
        const int16_t* fr = a.tdm + sig * a.in_stride + k * 12;
        a.out[sig * a.out_stride + k] =
            tdm_decim(tdm_mix(fr[0], fr[1], fr[2]), tdm_mix(fr[4], fr[5], fr[6]), tdm_mix(fr[8], fr[9], fr[10]));
    }
}

// ---- augment_audio_waveform ------------------------------------------------------------------------
struct AugArgs {
    const float* audio;   // [n][L] padded clips (extract_mfcc.py:157 pads before augmenting)
    float* out;           // [n][5][L]: original, speed 0.8, speed 1.2, volume 0.7, volume 1.3
    long long n;
    int L;                // 16000
    int len08, len12;     // int(L * 0.8), int(L * 1.2), computed on the host exactly as Python does
    float scale08, scale12;  // L / len (fp32), torch's area_pixel_compute_scale for size-given interpolation
    int vec_ok;           // 1: L % 4 == 0 and both buffers 16-byte aligned
};

// torch.nn.functional.interpolate(mode='linear', align_corners=False) source index (UpSample.h)
__device__ __forceinline__ float interp_linear(const float* x, int L, int dst, float scale) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    const int i0 = (int)s;
    const int i1 = i0 + (i0 < L - 1 ? 1 : 0);
    const float l1 = s - (float)i0, l0 = 1.f - l1;
    return l0 * x[i0] + l1 * x[i1];
}

// 4 consecutive samples per thread: one 16-byte load of the clip, five 16-byte streaming stores; the interpolation
// taps (two per output, source step 1.25 / 0.83) come through L1 from the same lines the neighbours load.
__global__ void __launch_bounds__(256) augment_kernel(const AugArgs a) {
    const int L4 = a.L / 4;                       // vector path requires L % 4 == 0 (host-checked), else scalar below
    if (a.vec_ok) {
        const long long total = a.n * L4;
        for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
            const long long c = it / L4;
            const int i = (int)(it - c * L4) * 4;
            const float* x = a.audio + c * a.L;
            float* o = a.out + c * 5 * a.L + i;
            const float4 v = *reinterpret_cast<const float4*>(x + i);
            float4 s08, s12;
            s08.x = i + 0 < a.len08 ? interp_linear(x, a.L, i + 0, a.scale08) : 0.f;
            s08.y = i + 1 < a.len08 ? interp_linear(x, a.L, i + 1, a.scale08) : 0.f;
            s08.z = i + 2 < a.len08 ? interp_linear(x, a.L, i + 2, a.scale08) : 0.f;
            s08.w = i + 3 < a.len08 ? interp_linear(x, a.L, i + 3, a.scale08) : 0.f;
            s12.x = i + 0 < a.len12 ? interp_linear(x, a.L, i + 0, a.scale12) : 0.f;
            s12.y = i + 1 < a.len12 ? interp_linear(x, a.L, i + 1, a.scale12) : 0.f;
            s12.z = i + 2 < a.len12 ? interp_linear(x, a.L, i + 2, a.scale12) : 0.f;
            s12.w = i + 3 < a.len12 ? interp_linear(x, a.L, i + 3, a.scale12) : 0.f;
            const float4 lo = make_float4(fminf(fmaxf(v.x * 0.7f, -1.f), 1.f), fminf(fmaxf(v.y * 0.7f, -1.f), 1.f),
                                          fminf(fmaxf(v.z * 0.7f, -1.f), 1.f), fminf(fmaxf(v.w * 0.7f, -1.f), 1.f));
            const float4 hi = make_float4(fminf(fmaxf(v.x * 1.3f, -1.f), 1.f), fminf(fmaxf(v.y * 1.3f, -1.f), 1.f),
                                          fminf(fmaxf(v.z * 1.3f, -1.f), 1.f), fminf(fmaxf(v.w * 1.3f, -1.f), 1.f));
            __stcs(reinterpret_cast<float4*>(o), v);
            __stcs(reinterpret_cast<float4*>(o + a.L), s08);
            __stcs(reinterpret_cast<float4*>(o + 2 * (long long)a.L), s12);
            __stcs(reinterpret_cast<float4*>(o + 3 * (long long)a.L), lo);
            __stcs(reinterpret_cast<float4*>(o + 4 * (long long)a.L), hi);
        }
        return;
    }
    const long long total = a.n * a.L;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const long long c = it / a.L;
        const int i = (int)(it - c * a.L);
        const float* x = a.audio + c * a.L;
        float* o = a.out + c * 5 * a.L + i;
        const float v = x[i];
        o[0] = v;
        // speed 0.8: 12800 interpolated samples then pad_audio's zero padding; speed 1.2: truncated to L
        o[a.L] = i < a.len08 ? interp_linear(x, a.L, i, a.scale08) : 0.f;
        o[2 * a.L] = i < a.len12 ? interp_linear(x, a.L, i, a.scale12) : 0.f;
        o[3 * a.L] = fminf(fmaxf(v * 0.7f, -1.f), 1.f);
        o[4 * a.L] = fminf(fmaxf(v * 1.3f, -1.f), 1.f);
    }
}

}  // namespace ww
