// Front-of-frontend DSP, sm_100a.
//
//   tdm_downmix_kernel  4-channel TDM 48 kHz int16 -> mono 16 kHz int16, the integer arithmetic of record_task
//                       (main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:103-121): bit-exact
//   augment_kernel      augment_audio_waveform (ml_models/src/extract_mfcc.py:90-121): the five deterministic
//                       variants of a padded clip (original, speed 0.8 / 1.2 by linear interpolation, volume 0.7 / 1.3)
//
// Both are streaming byte/element kernels bound by HBM: 26 B (24 in + 2 out) per 16 kHz sample for the down-mix,
// 4 B in + 20 B out per sample for the augmentation.  Loads are 16-byte vectors, fully coalesced; the grids are
// sized in multiples of the SM count and walk the work with a grid-stride loop.
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

__global__ void __launch_bounds__(256) tdm_downmix_kernel(const TdmArgs a) {
    // work item = 4 output samples = 12 TDM frames = 96 B = six 16-byte loads
    const long long quads = (a.n_out + 3) / 4;
    const long long total = a.n_signals * quads;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const long long sig = it / quads, qd = it - sig * quads;
        const int16_t* src = a.tdm + sig * a.in_stride + qd * 48;
        int16_t* dst = a.out + sig * a.out_stride + qd * 4;
        const long long left = a.n_out - qd * 4;
        if (a.vec_ok && left >= 4) {
            uint4 v[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = __ldcs(reinterpret_cast<const uint4*>(src) + i);  // streamed once
            int m[12];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                // uint4 = two frames: {L|ref, R|ch3, L|ref, R|ch3}
                m[2 * i] = tdm_mix(lo16(v[i].x), hi16(v[i].x), lo16(v[i].y));
                m[2 * i + 1] = tdm_mix(lo16(v[i].z), hi16(v[i].z), lo16(v[i].w));
            }
            const uint32_t o0 = (uint16_t)tdm_decim(m[0], m[1], m[2]), o1 = (uint16_t)tdm_decim(m[3], m[4], m[5]);
            const uint32_t o2 = (uint16_t)tdm_decim(m[6], m[7], m[8]), o3 = (uint16_t)tdm_decim(m[9], m[10], m[11]);
            *reinterpret_cast<uint2*>(dst) = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
        } else {
            for (int k = 0; k < 4 && k < left; ++k) {
                int m[3];
                for (int f = 0; f < 3; ++f) {
                    const int16_t* fr = src + (3 * k + f) * 4;
                    m[f] = tdm_mix(fr[0], fr[1], fr[2]);
                }
                dst[k] = tdm_decim(m[0], m[1], m[2]);
            }
        }
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

__global__ void __launch_bounds__(256) augment_kernel(const AugArgs a) {
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
