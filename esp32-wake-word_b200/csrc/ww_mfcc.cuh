// MFCC frontend kernel (sm_100a): PCM -> [13 x T] cepstra with no intermediate in HBM.
//
// Replaces, per clip / stream block, the reference chain
//   torchaudio.functional.preemphasis(x, 0.97)            ml_models/src/extract_mfcc.py:171
//   T.MFCC(16 kHz, 13, log_mels, n_fft 512, win 320, hop 256, 40 mels, hamming)   :137-148,172
// and, with the ESP table set, main/esp_mfcc/mfcc.c:431-527 (extract_mfcc).
//
// Work unit = one block of 32 consecutive frames of one signal (a 1 s clip is two blocks, 63 frames valid).
// The grid is persistent (two 8-warp CTAs per SM); each CTA walks over blocks and prefetches the PCM of its next
// block with TMA while it transforms the current one.
//   * the PCM span of the block is staged once into shared memory with a 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier); int16 PCM is kept as int16 in smem (16.5 KB per block)
//   * a warp transforms TWO frames at a time, 16 lanes each (the even lanes one frame, the odd lanes the other:
//     WW_LANE_INTERLEAVE below): the 512-point real FFT is a 256-point
//     complex FFT of the packed frame (z[m] = x[2m] + i x[2m+1]) done as radix-16 (registers) x
//     radix-16 (registers) with one transposition through a per-warp smem tile, followed by the
//     real-FFT split that yields two power bins per butterfly
//   * pre-emphasis and the window are folded into the load (w*(x[i] - 0.97*x[i-1])); only the 160 complex
//     points under the 320-tap window are loaded.  The block's PCM is staged as two halves whose smem
//     bases differ by 64 B mod 128, so the two frames of a warp (t and t+16) hit disjoint banks
//   * the real-FFT split takes its partner values Z[256-k] with 16 shuffles inside the frame's 16 lanes instead of parking Z in
//     shared memory: the kernel is co-limited by issue slots and shared-memory wavefronts, not by HBM
//   * the 32 power spectra of the block are parked in shared memory (33 KB); after a CTA barrier the mel
//     filterbank runs with lane <-> frame and warp <-> filter range as generated straight-line code
//     (ww_mel_py.inc) whose weights are FFMA immediates; log-mel rows go back to smem and the 40x13 DCT
//     runs the same way with its matrix in the kernel-parameter constant bank
//   * output is written once, coalesced along time
#pragma once
#include "ww_common.cuh"

// Table-traffic switches, A/B-measured on B200 (262 144 clips, tools/time_frontend.py, profiles/experiments/README.md):
//   window taps in 20 registers instead of 10 LDS.64 per frame pair                      26.6 -> 27.4 M clips/s (kept,
//                   no switch any more: the taps are read from global memory once and never staged in smem)
//   WW_TW2_COMPUTE  W512^(l16+16i) = W512^l16 * W32^i (immediates) instead of 8 LDS.64   +0.2 %  (off: noise)
//   WW_TW1_HALF     twiddles k1 >= 8 as W^(8 l16) * W^(l16 (k1-8)) instead of 4 LDS.128  -1 %    (off)
#ifndef WW_I2FP
#define WW_I2FP 0
#endif
#ifndef WW_TW2_COMPUTE
#define WW_TW2_COMPUTE 0
#endif
#ifndef WW_TW1_HALF
#define WW_TW1_HALF 0
#endif
// WW_TW1_REGS = N: the first N inter-pass twiddle entries of a lane in registers (the launch bound leaves a dozen free).
// Measured after the lane interleave (profiles/r2f_ab_tw1_regs.txt): N = 0 / 1 / 2 / 3 / 4 / 6 -> 33.70 / 33.25 / 33.19 /
// 33.50 / 32.51 / 32.58 M clips/s: ptxas pays for the longer live ranges with a worse schedule (off)
#ifndef WW_TW1_REGS
#define WW_TW1_REGS 0
#endif
// Round-2 switches (A/B on B200, profiles/experiments/README.md):
//   WW_IPRE     pre-emphasis in exact integer arithmetic, 100 x[i] - 97 x[i-1] by one IDP.2A per sample (|.| < 2^23, exact
//               in fp32), ONE int -> float conversion per sample on the ALU side (I2FP.F32.S32) instead of three
//               quarter-rate XU conversions (I2F.S16) per complex point; the 1/100 rides in the window taps
//   WW_PRESEL   the lane-0 special case of the real-FFT split is resolved on the SOURCE side of the shuffles
#ifndef WW_IPRE
#define WW_IPRE 1   // 31.47 -> 32.46 M clips/s (262 144 clips, alternating runs, gpurun_out/r2_ab_ipre.txt)
#endif
#ifndef WW_PRESEL
#define WW_PRESEL 0
#endif
// WW_LANE_INTERLEAVE: the two frames of a warp take the EVEN and the ODD lanes (frame = lane & 1, point index = lane >> 1)
// instead of the two half-warps.  Both frames read the same twiddles, and a 64-/128-bit shared load is served per
// half-warp: lanes i and i + 16 reading the same 16 bytes cost 4 wavefronts (the half-warps are not merged), lanes 2j and
// 2j + 1 reading the same 16 bytes cost 2 (tools/ubench/lds_dup.cu, profiles/r2e_ubench_lds_dup.txt: 4.00 against 2.11
// clocks per LDS.128, 1.75 against 0.98 per LDS.64).  The twiddle loads were 24.5 % of the kernel's shared-memory
// wavefronts; the two frames' exchange tiles are staggered by 64 B so that the transposition stays conflict-free.
#ifndef WW_LANE_INTERLEAVE
#define WW_LANE_INTERLEAVE 1
#endif
// WW_SPLIT_BARRIER (clip-shape instantiation): the CTA barrier at the end of a block becomes arrive ... wait on an
// mbarrier, and the edge-frame taps of the NEXT block (every block of a 1 s clip has one edge frame: 320 reflected,
// pre-emphasised taps, ~70 instructions per warp) are materialised between the two, i.e. while the warp would
// otherwise idle until the slowest warp of the CTA arrives (6 % of the stall samples).  The two tap slots alternate per
// iteration instead of per edge kind.
#ifndef WW_SPLIT_BARRIER
#define WW_SPLIT_BARRIER 1
#endif
// WW_STAGE_WARP: the warp whose lane 0 issues the TMA copies of the next block (~60 instructions executed by one lane,
// a serial chain that costs its warp several hundred clocks per block).  Per-warp arrival at the block barrier, measured
// with the -DWW_MFCC_STATS build and tools/mfcc_warp_slack.py (profiles/r2f_mfcc_warp_slack.txt): with warp 0 issuing
// them it arrives ~160 clocks after the mean of the eight and is the warp the CTA waits for.  Measured (1 048 576
// clips, alternating builds, profiles/r2f_ab_stage_warp.txt): warp 0 / 1 / 2 / 4: 33.74 / 33.93 / 33.93 / 33.80 M clips/s;
// the two halves issued by two warps (4 + 1, 1 + 2): 33.51 / 33.71 (two arrivals, the address arithmetic twice).
#ifndef WW_STAGE_WARP
#define WW_STAGE_WARP 1
#endif
// WW_FILL2_WARP: the 64 edge taps that are left over after one tap per thread are filled by this warp and the next one
// (warps 0 and 1 before: 34.00 M clips/s; 4 and 5: 33.78 M; 5 and 6: 34.05 M, profiles/r2g_ab_fill2.txt)
#ifndef WW_FILL2_WARP
#define WW_FILL2_WARP 5
#endif

namespace ww {

#ifdef WW_MFCC_STATS
// diagnostic build: per (CTA, warp) sum of clocks between a warp's arrival at the block barrier and the barrier's
// completion as that warp sees it (tools/mfcc_warp_slack.py)
__device__ unsigned long long g_mfcc_slack[1024 * 8];
#endif

// ---- table blob (one per feature mode, device memory, copied to smem by every CTA) ----------------
constexpr int TB_WIN_OFF = 0;                       // float2[160]  {w[2m], w[2m+1]} for packed point m = 48 + i
constexpr int TB_TW1_OFF = TB_WIN_OFF + 160 * 8;    // float4[8][16] {W256^(l*2j), W256^(l*(2j+1))}
constexpr int TB_TW2_OFF = TB_TW1_OFF + 128 * 16;   // float2[132]  W512^k, k = 0..128
constexpr int TB_MELW_OFF = TB_TW2_OFF + 132 * 8;   // float[640]   filterbank weights, filter-major
constexpr int TB_MELM_OFF = TB_MELW_OFF + 640 * 4;  // int4[40]     {start, len, off, bias bits}
constexpr int TB_BYTES = TB_MELM_OFF + 40 * 16;     // 8864
static_assert(TB_BYTES % 16 == 0, "table blob must be a multiple of 16 bytes");
constexpr int MEL_W_CAP = 640;

constexpr int MFCC_THREADS = 256;
constexpr int MFCC_WARPS = MFCC_THREADS / 32;
constexpr int EXCH_ROW_BYTES = 144;                    // 16 complex + 16 B pad: conflict-free LDS.128
constexpr int EXCH_FRAME_BYTES = 16 * EXCH_ROW_BYTES;  // 2304 (>= 257 complex for the natural-order pass)
// a warp's two exchange tiles: with interleaved lanes a half-warp touches both, so the second one sits 16 banks further
constexpr int EXCH_STAGGER = WW_LANE_INTERLEAVE ? 64 : 0;
constexpr int EXCH_WARP_BYTES = 2 * EXCH_FRAME_BYTES + EXCH_STAGGER;
// Rows of the power-spectrum and log-mel buffers (lane <-> frame in the mel / DCT stages).  The generated mel / DCT
// code reads its row with 16-byte loads: rows are 16-byte aligned with a stride of 4 (mod 32) words, which makes the
// eight lanes of a quarter-warp hit disjoint banks, and the rows of frames 16..31 are shifted by another 16 words so
// that the two half-warps of an FFT warp (frames t and t + 16, scalar stores along k) stay on disjoint banks too.
// The table-driven fallback reads scalars and keeps odd strides.
template <int MEL>
struct MfccRows {
    static constexpr bool VEC = MEL != 0;
    static constexpr int P_STRIDE = VEC ? 260 : 257;
    static constexpr int P_HALF_SHIFT = VEC ? 16 : 0;
    static constexpr int P_FLOATS = 32 * P_STRIDE + P_HALF_SHIFT;
    static constexpr int LM_STRIDE = VEC ? 44 : 41;
    static __device__ __forceinline__ int p_row(int fl) { return fl * P_STRIDE + (fl >> 4) * P_HALF_SHIFT; }
};

struct MfccArgs {
    const void* pcm;          // [n_signals][sig_stride] samples
    long long sig_stride;     // samples between signals
    int n_samples;            // valid samples per signal
    int n_frames;             // frames per signal
    int blocks_per_sig;       // ceil(n_frames / FRAMES)
    long long n_blocks;       // n_signals * blocks_per_sig
    float* out;               // out[sig*out_sig_stride + coef*out_coef_stride + frame*out_frame_stride]
    long long out_sig_stride;
    long long out_coef_stride;
    long long out_frame_stride;
    const uint4* tables;      // TB_BYTES blob
    int origin_off;           // sample index of FFT-frame point n=0 relative to 256*t: -256 (PY), -96 (ESP)
    int reflect;              // 1: reflect-pad at the signal ends (torch.stft center=True)
    int use_bulk;             // 1: TMA bulk copy is legal for this launch (alignment checked on host)
    float pscale;             // power scale folded after the mel sum (0.25 * input scale^2 * mode scale)
    float log_floor;          // lm = log(max(mel + bias, floor) + offset)
    float log_offset;
    float preemph;            // 0.97
    float pm1[2];             // {1, -1}: a run-time value so that it stays in one register pair (see the kernel)
    float dct[WW_N_MELS * WW_N_MFCC];  // row-major [40][13]; lives in the parameter constant bank
};

template <typename TIN>
__device__ __forceinline__ float pcm_to_float(TIN v);
template <>
__device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return static_cast<float>(v); }
template <>
__device__ __forceinline__ float pcm_to_float<float>(float v) { return v; }

// ---- 16-point complex FFT in registers (forward, natural order in and out), packed f32x2 arithmetic ---
__device__ __forceinline__ void dft4(cpx& p0, cpx& p1, cpx& p2, cpx& p3) {
    const cpx t0 = p_add(p0, p2), t1 = p_sub(p0, p2), t2 = p_add(p1, p3), d = p_sub(p1, p3);
    p0 = p_add(t0, t2);
    p1 = p_add_mi(t1, d);
    p2 = p_sub(t0, t2);
    p3 = p_sub_mi(t1, d);
}

// DFT4 with compile-time knowledge of zero inputs (Z0..Z3): skips the additions with 0
template <bool Z0, bool Z1, bool Z2, bool Z3>
__device__ __forceinline__ void dft4z(cpx& p0, cpx& p1, cpx& p2, cpx& p3) {
    static_assert(!Z1 && !Z2, "only the outer inputs are ever known to be zero");
    const cpx t0 = Z0 ? p2 : p_add(p0, p2);
    const cpx t1 = Z0 ? p_neg(p2) : p_sub(p0, p2);
    const cpx t2 = Z3 ? p1 : p_add(p1, p3);
    const cpx d = Z3 ? p1 : p_sub(p1, p3);
    p0 = p_add(t0, t2);
    p1 = p_add_mi(t1, d);
    p2 = p_sub(t0, t2);
    p3 = p_sub_mi(t1, d);
}

// WINDOWED: inputs v[0..2] and v[13..15] are known to be zero (only 160 of the 256 packed points lie
// under the 320-tap window), which prunes a third of the first butterfly stage.
// 80 packed instructions (46 FADD2 + 26 FFMA2 + 8 FMUL2) against 174 scalar ones.
template <bool WINDOWED>
__device__ __forceinline__ void fft16(cpx (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
    constexpr float R2 = 0.70710678118654752f;  // sqrt(1/2)
    // n = 4a + b, k = c + 4d.  Step 1: DFT4 over a for each b  -> v[4c + b] = Y[b][c]
    if constexpr (WINDOWED) {
        dft4z<true, false, false, false>(v[0], v[4], v[8], v[12]);
        dft4z<true, false, false, true>(v[1], v[5], v[9], v[13]);
        dft4z<true, false, false, true>(v[2], v[6], v[10], v[14]);
        dft4z<false, false, false, true>(v[3], v[7], v[11], v[15]);
    } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    }
    // Step 2: twiddle W16^(b*c)
    v[4 * 1 + 1] = p_cmul(v[4 * 1 + 1], C1, -S1);   // W^1
    v[4 * 1 + 2] = p_cmul(v[4 * 1 + 2], R2, -R2);   // W^2
    v[4 * 1 + 3] = p_cmul(v[4 * 1 + 3], S1, -C1);   // W^3
    v[4 * 2 + 1] = p_cmul(v[4 * 2 + 1], R2, -R2);   // W^2
    v[4 * 2 + 2] = p_mul_mi(v[4 * 2 + 2]);          // W^4 = -i
    v[4 * 2 + 3] = p_cmul(v[4 * 2 + 3], -R2, -R2);  // W^6
    v[4 * 3 + 1] = p_cmul(v[4 * 3 + 1], S1, -C1);   // W^3
    v[4 * 3 + 2] = p_cmul(v[4 * 3 + 2], -R2, -R2);  // W^6
    v[4 * 3 + 3] = p_cmul(v[4 * 3 + 3], -C1, S1);   // W^9
    // Step 3: DFT4 over b for each c -> v[4c + d] = X[c + 4d]
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // natural order: X[k] = v[4*(k&3) + (k>>2)]  (register renaming only)
    cpx r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = v[4 * (k & 3) + (k >> 2)];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = r[k];
}

// ---- generic (edge-frame) sample fetch: reflect + pre-emphasis, sample index s relative to the signal
template <typename TIN>
__device__ __forceinline__ float emph_sample(const TIN* spcm, int lo, int s, int L, int reflect, float pre) {
    if (reflect) {
        if (s < 0) s = -s;
        if (s >= L) s = 2 * (L - 1) - s;
    }
    if (s < 0 || s >= L) return 0.f;
    float x = pcm_to_float<TIN>(spcm[s - lo]);
    if (s > 0) x = __fsub_rn(x, __fmul_rn(pre, pcm_to_float<TIN>(spcm[s - 1 - lo])));
    return x;
}

constexpr int MFCC_FRAMES = 32;   // frames per CTA (lane <-> frame in the mel / DCT phases)
constexpr int TB_BYTES_PY = TB_MELW_OFF;  // the PY path needs no mel tables in smem

// MEL: 0 = table-driven filterbank / DCT from the blob (any contiguous-support filter set), 1 = generated PY code
// (ww_mel_py.inc), 2 = generated ESP code (ww_mel_esp.inc): weights as FFMA immediates, no mel tables in smem
constexpr int MEL_TABLE = 0, MEL_PY = 1, MEL_ESP = 2;

template <typename TIN, int MEL>
struct MfccSmem {
    static constexpr int FRAMES = MFCC_FRAMES;
    // the block is staged as two halves of 16 frames: 15*256 + 320 taps + 8 lead samples each
    static constexpr int HALF_SAMPLES = 15 * WW_HOP + 328;
    static constexpr int HALF_RAW = HALF_SAMPLES * (int)sizeof(TIN);
    // second half starts 64 B (mod 128) after the first: frames t and t+16 then read disjoint banks
    static constexpr int HALF_STRIDE = HALF_RAW + ((64 - HALF_RAW % 128) + 128) % 128;
    static constexpr int PCM_BYTES = HALF_STRIDE + ((HALF_RAW + 15) / 16) * 16;
    static_assert(HALF_STRIDE % 16 == 0 && HALF_STRIDE % 128 == 64, "half-warp bank stagger");
    // int16 PCM on the PY path is double-buffered (the next block is prefetched by TMA while this one is
    // transformed); the other variants are single-buffered to keep two CTAs per SM
    static constexpr int PCM_BUFS = (sizeof(TIN) == 2 && MEL != MEL_TABLE) ? 2 : 1;
    // the window taps (head of the blob) live in registers and are never staged: smem holds blob[TB_TW1_OFF ..)
    static constexpr int TAB_SKIP = TB_TW1_OFF;
    static constexpr int TAB_BYTES = (MEL != MEL_TABLE ? TB_BYTES_PY : TB_BYTES) - TAB_SKIP;
    static constexpr int OFF_BAR = 0;   // five mbarriers
    static constexpr int OFF_TAB = 48;
    static constexpr int OFF_PCM = OFF_TAB + TAB_BYTES;
    static constexpr int OFF_EXCH = OFF_PCM + PCM_BUFS * PCM_BYTES;
    static constexpr int OFF_P = OFF_EXCH + MFCC_WARPS * EXCH_WARP_BYTES;
    static constexpr int OFF_LM = OFF_P + MfccRows<MEL>::P_FLOATS * 4;
    static constexpr int OFF_EDGE = OFF_LM + FRAMES * MfccRows<MEL>::LM_STRIDE * 4;   // 2 x 320 pre-emphasised edge-frame samples
    static constexpr int TOTAL = OFF_EDGE + 2 * WW_WIN * 4;
    static_assert(OFF_PCM % 16 == 0 && OFF_EXCH % 16 == 0 && OFF_P % 16 == 0 && OFF_LM % 16 == 0, "smem alignment");
    // two CTAs per SM need TOTAL <= 115712 B; only the (float PCM, table-driven mel) variant exceeds it
    static_assert(TOTAL <= 115712 || (MEL == MEL_TABLE && sizeof(TIN) == 4), "two CTAs per SM must fit");
};

#include "ww_mel_py.inc"
#include "ww_mel_esp.inc"
static_assert(WW_MEL_PY_GROUPS == MFCC_WARPS && WW_MEL_ESP_GROUPS == MFCC_WARPS, "one generated mel group per warp");

// DCT of one frame for the coefficient subset {G0, G0+G, ...}: weights are immediate constant operands.
template <int G, int G0>
__device__ __forceinline__ void dct_store(const MfccArgs& a, const float* lm_row, float* outp) {
    constexpr int NQ = (WW_N_MFCC - G0 + G - 1) / G;
    float acc[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < WW_N_MELS; ++j) {
        const float l = lm_row[j];
#pragma unroll
        for (int i = 0; i < NQ; ++i) acc[i] = fmaf(l, a.dct[j * WW_N_MFCC + G0 + G * i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) outp[(long long)(G0 + G * i) * a.out_coef_stride] = acc[i];
}

template <int G, int G0>
struct DctDispatch {
    static __device__ __forceinline__ void run(int g, const MfccArgs& a, const float* lm_row, float* outp) {
        if (g == G0) dct_store<G, G0>(a, lm_row, outp);
        else DctDispatch<G, G0 + 1>::run(g, a, lm_row, outp);
    }
};
template <int G>
struct DctDispatch<G, G> {
    static __device__ __forceinline__ void run(int, const MfccArgs&, const float*, float*) {}
};

// CLIP = true is the same kernel with the launch shape of the headline workload frozen at compile time (whole 1 s
// clips of the PY feature mode: 16 000 contiguous samples, 63 frames, two blocks per clip, reflect padding, TMA-legal
// alignment, [B][13][63] output): the per-block bookkeeping (block -> signal mapping, staging spans, edge-frame tests,
// output addressing) that every warp repeats for every block folds into constants.  The arithmetic is untouched, so
// both instantiations produce identical bits; the host picks CLIP when the launch arguments match (launch_mfcc_ex).
constexpr int CLIP_SAMPLES = 16000, CLIP_FRAMES = 63, CLIP_ORIGIN = -256;

// Where one frontend pipeline (256 threads, MfccSmem bytes of shared memory) runs.  The stand-alone kernel runs one per
// CTA (two CTAs per SM); the fused clip kernel (ww_fused.cuh) runs two per 512-thread CTA, each with its own named
// barrier, and hands the features to the CNN role of the same launch through an L2-resident ring (FUSED: out_slot /
// free_probe / wait_free / publish).
// WW_PCM_EVICT_FIRST: the TMA loads of the PCM carry an evict-first L2 policy (every sample is read exactly once; without
// the hint the 32 KB/clip input stream pushes the tables, the output lines and the next kernel's input out of L2)
#ifndef WW_PCM_EVICT_FIRST
#define WW_PCM_EVICT_FIRST 1
#endif
struct MfccSolo {
    static constexpr bool FUSED = false;
    uint64_t pol_ = 0;
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ long long first() const { return blockIdx.x; }
    __device__ __forceinline__ long long stride() const { return gridDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ long long out_slot(long long sig) const { return sig; }
    __device__ __forceinline__ int free_probe(long long) const { return 0; }
    __device__ __forceinline__ void wait_free(long long, int) const {}
    __device__ __forceinline__ void publish(long long) const {}
    __device__ __forceinline__ uint64_t pcm_policy() const { return pol_; }
};

template <typename TIN, int MEL, bool CLIP, class PIPE>
__device__ __forceinline__ void mfcc_body(const MfccArgs& a, unsigned char* smem, const PIPE pipe) {
    using SM = MfccSmem<TIN, MEL>;
    constexpr int FRAMES = MFCC_FRAMES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    using ROWS = MfccRows<MEL>;
    const unsigned char* tab = smem + SM::OFF_TAB - SM::TAB_SKIP;  // tab + TB_x_OFF for x = TW1, TW2, MELW, MELM
    float* pw = reinterpret_cast<float*>(smem + SM::OFF_P);
    float* lm = reinterpret_cast<float*>(smem + SM::OFF_LM);

    // warp index out of a shuffle (known warp-uniform to the compiler, ww_common.cuh): 32.36 -> 32.49 M clips/s, same bits
    const int tid = pipe.tid(), warp = warp_index_uniform(tid), lane = tid & 31;
    // which of the warp's two frames this lane works on, and its point index inside the frame
    const int half = WW_LANE_INTERLEAVE ? (lane & 1) : (lane >> 4), l16 = WW_LANE_INTERLEAVE ? (lane >> 1) : (lane & 15);

    const int L = CLIP ? CLIP_SAMPLES : a.n_samples;
    const int n_frames = CLIP ? CLIP_FRAMES : a.n_frames;
    const int origin_off = CLIP ? CLIP_ORIGIN : a.origin_off;
    const bool use_bulk = CLIP ? true : a.use_bulk != 0;
    const long long sig_stride = CLIP ? (long long)CLIP_SAMPLES : a.sig_stride;
    const long long out_sig_stride = CLIP ? (long long)(WW_N_MFCC * CLIP_FRAMES) : a.out_sig_stride;
    const long long out_coef_stride = CLIP ? (long long)CLIP_FRAMES : a.out_coef_stride;
    const long long out_frame_stride = CLIP ? 1LL : a.out_frame_stride;
    const int reflect = CLIP ? 1 : a.reflect;
    constexpr int NBUF = SM::PCM_BUFS;
    constexpr int kStageTid = 32 * WW_STAGE_WARP;
    const bool stager = tid == kStageTid;
    uint64_t* bars = bar;  // one mbarrier per PCM buffer

    // half h of a block holds frames 16h..16h+15; smem sample index = s - org[h]
    // one thread: TMA both halves of block (signal sg, block-in-signal bi) into buffer `buf`
    auto stage_block = [&](long long sg, int bi, int buf) {
        const int bt0 = bi * FRAMES;
        const TIN* gs = reinterpret_cast<const TIN*>(a.pcm) + sg * sig_stride;
        int lo[2], n[2];
        for (int h = 0; h < 2; ++h) {
            const int org = WW_HOP * (bt0 + 16 * h) + origin_off + 88;
            lo[h] = org < 0 ? 0 : org;
            int hi = org + SM::HALF_SAMPLES;
            hi = hi > L ? L : hi;
            n[h] = hi > lo[h] ? hi - lo[h] : 0;
        }
        // byte counts are multiples of 16 by construction when use_bulk is set (host-checked)
        mbar_expect_tx(&bars[buf], (uint32_t)((n[0] + n[1]) * (int)sizeof(TIN)));
        for (int h = 0; h < 2; ++h) {
            const int org = WW_HOP * (bt0 + 16 * h) + origin_off + 88;
            if (n[h]) {
                unsigned char* dst = smem + SM::OFF_PCM + buf * SM::PCM_BYTES + h * SM::HALF_STRIDE + (lo[h] - org) * (int)sizeof(TIN);
                if ((PIPE::FUSED || WW_PCM_EVICT_FIRST) && pipe.pcm_policy() != 0)   // the PCM is read once
                    bulk_g2s_hint(dst, gs + lo[h], (uint32_t)(n[h] * (int)sizeof(TIN)), &bars[buf], pipe.pcm_policy());
                else
                    bulk_g2s(dst, gs + lo[h], (uint32_t)(n[h] * (int)sizeof(TIN)), &bars[buf]);
            }
        }
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], MFCC_WARPS);  // every warp has finished its mel stage (P free, LM complete)
        mbar_init(&bars[3], MFCC_WARPS);  // every warp has written its share of the edge taps
        mbar_init(&bars[4], MFCC_WARPS);  // WW_SPLIT_BARRIER: every warp has finished the block
        mbar_fence_init();
    }
    // tables -> smem (once per persistent CTA)
    {
        uint4* dst = reinterpret_cast<uint4*>(smem + SM::OFF_TAB);
        for (int i = tid; i < SM::TAB_BYTES / 16; i += MFCC_THREADS) dst[i] = __ldg(a.tables + SM::TAB_SKIP / 16 + i);
    }
    pipe.sync();
    const long long first = pipe.first(), stride = pipe.stride();
    // (signal, block-in-signal) of the current block, advanced incrementally: no 64-bit division in the loop
    const int bps = CLIP ? (CLIP_FRAMES + FRAMES - 1) / FRAMES : a.blocks_per_sig;
    const int stride_q = (int)stride / bps, stride_r = (int)stride % bps;
    long long sig_cur = first / bps;
    int bi_cur = (int)(first - sig_cur * bps);
    if (use_bulk && stager && first < a.n_blocks) stage_block(sig_cur, bi_cur, 0);

    float* edge = reinterpret_cast<float*>(smem + SM::OFF_EDGE);
    const int t_tail = (L - origin_off - 416) / WW_HOP + 1;  // first frame whose taps run past the signal end
    // WW_SPLIT_BARRIER: the one edge frame of a clip block (frame 0 of block 0, the tail frame of block 1), filled a block
    // ahead into tap slot `slot` from the PCM buffer the block was staged into
    constexpr bool EARLY_EDGE = WW_SPLIT_BARRIER && CLIP && NBUF == 2;   // needs the next block's PCM staged ahead
    auto fill_edge_ahead = [&](int bi, const unsigned char* pcm_b, int slot) {
        const int bt0 = bi * MFCC_FRAMES;
        const int te = bi == 0 ? 0 : t_tail;
        const int hh = (te - bt0) >> 4;
        const TIN* sp = reinterpret_cast<const TIN*>(pcm_b + hh * SM::HALF_STRIDE);
        const int s0 = WW_HOP * te + origin_off + 96, lo_h = WW_HOP * (bt0 + 16 * hh) + origin_off + 88;
        constexpr float es = (WW_IPRE && sizeof(TIN) == 2) ? 100.f : 1.f;
        static_assert(WW_WIN > MFCC_THREADS && WW_WIN - MFCC_THREADS <= 64, "one tap per thread and 64 left over");
        edge[slot * WW_WIN + tid] = es * emph_sample<TIN>(sp, lo_h, s0 + tid, L, reflect, a.preemph);
        // the 64 taps left over go to the two warps that reach the barrier first (WW_FILL2_WARP), not to warps 0 and 1
        const int j2 = MFCC_THREADS + tid - 32 * WW_FILL2_WARP;
        if (tid >= 32 * WW_FILL2_WARP && j2 < WW_WIN)
            edge[slot * WW_WIN + j2] = es * emph_sample<TIN>(sp, lo_h, s0 + j2, L, reflect, a.preemph);
    };
    if constexpr (EARLY_EDGE) {
        if (first < a.n_blocks) {
            mbar_wait(&bars[0], 0);
            fill_edge_ahead(bi_cur, smem + SM::OFF_PCM, 0);
        }
        pipe.sync();
    }

    const float4* s_tw1 = reinterpret_cast<const float4*>(tab + TB_TW1_OFF);
    const float2* s_tw2 = reinterpret_cast<const float2*>(tab + TB_TW2_OFF);

    unsigned char* exch = smem + SM::OFF_EXCH + warp * EXCH_WARP_BYTES + half * (EXCH_FRAME_BYTES + EXCH_STAGGER);

    // Per-thread constants that depend on l16 only can live in registers for the whole kernel.  Both frames of a warp
    // need the same table entries; with the frames on adjacent lanes (WW_LANE_INTERLEAVE) a 64-/128-bit table read is
    // merged inside each half-warp and costs half the wavefronts it cost with a frame per half-warp (the kernel runs
    // the shared-memory pipe at ~2/3 of its peak):
    //   window: 10 packed taps (20 registers)                                              [on]
    //   tw2:    W512^(l16 + 16 i) = W512^l16 * W32^i with W32^i as immediates              [measured, off]
    //   tw1:    W256^(l16 k1), k1 >= 8, = W256^(8 l16) * W256^(l16 (k1 - 8))               [measured, off]
    cpx wreg[10];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        wreg[i].v = __ldg(reinterpret_cast<const unsigned long long*>(a.tables) + TB_WIN_OFF / 8 + 16 * i + l16);
    // WW_IPRE (int16 PCM): samples reach the window as 100 x[i] - 97 x[i-1]; the taps carry the 1/100
    constexpr bool IPRE = WW_IPRE && sizeof(TIN) == 2;
    constexpr float emph_scale = IPRE ? 100.f : 1.f;
    if constexpr (IPRE) {
#pragma unroll
        for (int i = 0; i < 10; ++i) wreg[i] = p_mul(wreg[i], cpk(0.01f, 0.01f));
    }
#if WW_TW2_COMPUTE
    const float2 tw2_0 = s_tw2[l16];
#endif
    // {1, -1} comes from the launch arguments so that it lives in one register pair for the whole kernel: as a
    // compile-time constant it is re-materialised from a uniform register with two MOVs in front of every packed
    // instruction whose other two operands are broadcast scalars
    const cpx pm1 = cpk(a.pm1[0], a.pm1[1]);
#if WW_TW1_HALF
    const cpx tw1_8 = cpk(s_tw1[16 * 4 + l16].x, s_tw1[16 * 4 + l16].y);
#endif
    float4 tw1_r[WW_TW1_REGS > 0 ? WW_TW1_REGS : 1];
#pragma unroll
    for (int j = 0; j < WW_TW1_REGS; ++j) tw1_r[j] = __ldg(reinterpret_cast<const float4*>(a.tables) + TB_TW1_OFF / 16 + 16 * j + l16);

    // Per-CTA software pipeline over this CTA's blocks (one CTA-wide barrier per block):
    //   iteration k:  mel(k-1) | stage PCM(k) + edge taps | FFT(k) first pass | DCT(k-1) | FFT(k) second pass | barrier
    // mel(k-1) and DCT(k-1) run between the FFT passes without waiting: P(k-1) was completed by the barrier of
    // iteration k-1, and the only cross-warp conditions -- "every warp is done reading P(k-1)" before P(k) is
    // written, "every warp has written its log-mel rows" before the DCT, "edge taps are in place" -- are mbarriers
    // that are normally already complete when they are tested.
    //   bars[0..1] PCM buffers (TMA), bars[2] mel done (8 warps), bars[3] edge taps done (8 warps)
    constexpr unsigned kDctWarps = MEL == MEL_PY ? WW_DCT_PY_WARP_MASK : MEL == MEL_ESP ? WW_DCT_ESP_WARP_MASK : 0xffu;
    long long iter = 0;
    bool prev_valid = false;
    long long prev_sig = 0;
    bool pub_valid = false;      // FUSED: the block before the previous one has its features stored
    // FUSED: the hand-over (publish / free_probe / wait_free) is done by single lanes from inline asm.  Fed from sig_cur /
    // prev_sig it dragged the whole block -> clip bookkeeping off the uniform datapath (+12 % instructions); it gets its
    // own per-thread copy of the clip index instead (clip of iteration k = v_sig, advanced by stride_q: the fused grid
    // has an even stride, so stride_r = 0), hidden from the optimiser so that the two chains are not merged again.
    long long v_sig = 0;
    if constexpr (PIPE::FUSED) {
        v_sig = sig_cur;
        asm volatile("" : "+l"(v_sig));
    }
    int prev_t0 = 0;
    uint32_t mel_uses = 0, edge_uses = 0;
#pragma unroll 1
    for (long long blk_id = first;; blk_id += stride, ++iter) {
        const bool have = blk_id < a.n_blocks;
        if (!have && !prev_valid) break;
        bool mel_pending = false;
        // programmatic dependent launch: the first block was staged and transformed while the previous kernel of the
        // stream (the CNN launch that still reads the feature buffer this launch overwrites) was finishing; the first
        // output store happens in this iteration's DCT
        if (!PIPE::FUSED && iter == 1) pdl_wait();
        // FUSED: the block stored by the previous iteration's DCT is complete (the CTA barrier ordered its stores before
        // this point): one thread hands it to the CNN role.  The probe of the ring slot this iteration's DCT will write
        // is issued here, long before its result is needed.
        int free_seen = 0;
        if constexpr (PIPE::FUSED) {
            if (pub_valid && tid == MFCC_THREADS - 32) pipe.publish(v_sig - 2 * stride_q);
            // every lane of every warp reads the same word (one broadcast transaction per warp): a probe by the DCT
            // warps' lane 0 alone put a divergent definition in front of the mel stage and cost the kernel its
            // uniform datapath
            if (prev_valid) free_seen = pipe.free_probe(v_sig - stride_q);
        }

        // ---- mel + log of the previous block -> LM
        if (prev_valid) {
            // mel + log: lane <-> frame, warp <-> filter range
            {
                const float pscale = a.pscale, log_offset = a.log_offset;
                const float* prow = pw + ROWS::p_row(lane);
                float* lrow = lm + lane * ROWS::LM_STRIDE;
                if constexpr (MEL == MEL_PY) {
                    switch (warp) {
                        case 0: mel_py_group<0>(prow, lrow, pscale, log_offset); break;
                        case 1: mel_py_group<1>(prow, lrow, pscale, log_offset); break;
                        case 2: mel_py_group<2>(prow, lrow, pscale, log_offset); break;
                        case 3: mel_py_group<3>(prow, lrow, pscale, log_offset); break;
                        case 4: mel_py_group<4>(prow, lrow, pscale, log_offset); break;
                        case 5: mel_py_group<5>(prow, lrow, pscale, log_offset); break;
                        case 6: mel_py_group<6>(prow, lrow, pscale, log_offset); break;
                        default: mel_py_group<7>(prow, lrow, pscale, log_offset); break;
                    }
                } else if constexpr (MEL == MEL_ESP) {
                    const float log_floor = a.log_floor;
                    switch (warp) {
                        case 0: mel_esp_group<0>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 1: mel_esp_group<1>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 2: mel_esp_group<2>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 3: mel_esp_group<3>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 4: mel_esp_group<4>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 5: mel_esp_group<5>(prow, lrow, pscale, log_floor, log_offset); break;
                        case 6: mel_esp_group<6>(prow, lrow, pscale, log_floor, log_offset); break;
                        default: mel_esp_group<7>(prow, lrow, pscale, log_floor, log_offset); break;
                    }
                } else {
                    // table-driven filterbank (weights broadcast from smem): any contiguous-support filter set
                    const float* s_melw = reinterpret_cast<const float*>(tab + TB_MELW_OFF);
                    const int4* s_melm = reinterpret_cast<const int4*>(tab + TB_MELM_OFF);
                    const float log_floor = a.log_floor;
                    for (int j = warp; j < WW_N_MELS; j += MFCC_WARPS) {
                        const int4 m = s_melm[j];
                        const float* pp = prow + m.x;
                        const float* ww_ = s_melw + m.z;
                        float acc = 0.f;
                        for (int i = 0; i < m.y; ++i) acc = fmaf(pp[i], ww_[i], acc);
                        const float e = fmaf(acc, pscale, __int_as_float(m.w));
                        lrow[j] = __logf(fmaxf(e, log_floor) + log_offset);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[2]);
            mel_pending = true;
        }

        long long sig = 0;
        int t0 = 0, org0 = 0, org1 = 0;
        unsigned char* pcm_buf = smem + SM::OFF_PCM;
        bool block_has_edge = false;
        int it_first = 0;
        // the block after this one (same CTA)
        long long sig_nxt = sig_cur + stride_q;
        int bi_nxt = bi_cur + stride_r;
        if (bi_nxt >= bps) {
            bi_nxt -= bps;
            ++sig_nxt;
        }
        if (have) {
            sig = sig_cur;
            t0 = bi_cur * FRAMES;
            org0 = WW_HOP * t0 + origin_off + 88;  // 8 samples ahead of the first window tap (multiple of 8)
            org1 = org0 + 16 * WW_HOP;
            const int buf = NBUF == 2 ? (int)(iter & 1) : 0;
            pcm_buf = smem + SM::OFF_PCM + buf * SM::PCM_BYTES;
            if (use_bulk) {
                if (NBUF == 2) {
                    // the other buffer was last read in the FFT passes of the previous block (before its barrier)
                    if (stager && blk_id + stride < a.n_blocks) stage_block(sig_nxt, bi_nxt, buf ^ 1);
                    mbar_wait(&bars[buf], (uint32_t)((iter >> 1) & 1));
                } else {
                    if (iter > 0 && stager) stage_block(sig_cur, bi_cur, 0);
                    mbar_wait(&bars[0], (uint32_t)(iter & 1));
                }
            } else {
                // unaligned / odd-length signals: cooperative copy instead of TMA
                const TIN* gsig = reinterpret_cast<const TIN*>(a.pcm) + sig * sig_stride;
                for (int h = 0; h < 2; ++h) {
                    const int org = h ? org1 : org0;
                    const int lo_h = org < 0 ? 0 : org;
                    int hi_h = org + SM::HALF_SAMPLES;
                    hi_h = hi_h > L ? L : hi_h;
                    TIN* dst = reinterpret_cast<TIN*>(pcm_buf + h * SM::HALF_STRIDE) + (lo_h - org);
                    for (int i = tid; i < hi_h - lo_h; i += MFCC_THREADS) dst[i] = gsig[lo_h + i];
                }
                pipe.sync();
            }

            // Edge frames (t = 0: reflected / no previous sample; the last frame: reflected tail) cannot use the
            // packed fast path.  Their 320 pre-emphasised taps are materialised once, cooperatively, so that no
            // warp of the block is slower than the others; the frames that need them are scheduled in the second
            // FFT pass, by when the fill (signalled through bars[3]) has long completed.
            const bool has0 = (t0 == 0) && (origin_off + 95 < 0);
            const bool has1 = (t_tail >= t0) && (t_tail < t0 + FRAMES) && (t_tail < n_frames) && (t_tail > 0 || !has0);
            block_has_edge = has0 || has1;
            if (block_has_edge && EARLY_EDGE) {
                const int it_edge = has0 ? 0 : (((t_tail - t0) & 15) >> 3);
                it_first = it_edge ^ 1;
            } else if (block_has_edge) {
                const float pre = a.preemph;
    #pragma unroll 1
                for (int slot = has0 ? 0 : 1; slot <= (has1 ? 1 : 0); ++slot) {
                    const int te = slot ? t_tail : 0;
                    const int hh = (te - t0) >> 4;
                    const TIN* sp = reinterpret_cast<const TIN*>(pcm_buf + hh * SM::HALF_STRIDE);
                    const int s0 = WW_HOP * te + origin_off + 96, lo_h = hh ? org1 : org0;
                    for (int j = tid; j < WW_WIN; j += MFCC_THREADS)
                        edge[slot * WW_WIN + j] = emph_scale * emph_sample<TIN>(sp, lo_h, s0 + j, L, reflect, pre);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[3]);
                const int it_edge = has0 ? 0 : (((t_tail - t0) & 15) >> 3);
                it_first = it_edge ^ 1;
            }
        }

        constexpr int ITERS = FRAMES / (2 * MFCC_WARPS);
        static_assert(ITERS == 2, "the pipeline interleaves the DCT between exactly two FFT passes");
#pragma unroll 1
        for (int trip = 0; trip < ITERS; ++trip) {
            if (trip == 1 && prev_valid) {
                if (mel_pending) {  // no FFT pass waited for it (drain iteration)
                    mbar_wait(&bars[2], mel_uses & 1);
                    mel_pending = false;
                }
                // DCT: lane <-> frame, four of the warps <-> three or four coefficients each (the generated code's
                // WW_DCT_*_WARP_MASK; every DCT warp reads the whole log-mel row); stores coalesced along time
                constexpr unsigned dct_warps = kDctWarps;
                if ((dct_warps >> warp) & 1u) {
                    if constexpr (PIPE::FUSED) pipe.wait_free(v_sig - stride_q, free_seen);
                    const int t = prev_t0 + lane;
                    if (t < n_frames) {
                        float* outp = a.out + pipe.out_slot(prev_sig) * out_sig_stride + (long long)t * out_frame_stride;
                        const float* lrow = lm + lane * ROWS::LM_STRIDE;
                        if constexpr (MEL == MEL_PY) {
                            const long long cs = out_coef_stride;
                            switch (warp) {
                                case 0: dct_py_group<0>(lrow, outp, cs); break;
                                case 1: dct_py_group<1>(lrow, outp, cs); break;
                                case 2: dct_py_group<2>(lrow, outp, cs); break;
                                case 3: dct_py_group<3>(lrow, outp, cs); break;
                                case 4: dct_py_group<4>(lrow, outp, cs); break;
                                case 5: dct_py_group<5>(lrow, outp, cs); break;
                                case 6: dct_py_group<6>(lrow, outp, cs); break;
                                default: dct_py_group<7>(lrow, outp, cs); break;
                            }
                        } else if constexpr (MEL == MEL_ESP) {
                            const long long cs = out_coef_stride;
                            switch (warp) {
                                case 0: dct_esp_group<0>(lrow, outp, cs); break;
                                case 1: dct_esp_group<1>(lrow, outp, cs); break;
                                case 2: dct_esp_group<2>(lrow, outp, cs); break;
                                case 3: dct_esp_group<3>(lrow, outp, cs); break;
                                case 4: dct_esp_group<4>(lrow, outp, cs); break;
                                case 5: dct_esp_group<5>(lrow, outp, cs); break;
                                case 6: dct_esp_group<6>(lrow, outp, cs); break;
                                default: dct_esp_group<7>(lrow, outp, cs); break;
                            }
                        } else {
                            DctDispatch<MFCC_WARPS, 0>::run(warp, a, lrow, outp);
                        }
                    }
                }
            }
            if (!have) continue;
            const int it = trip ^ it_first;
            // the two half-warps take frames 16 apart: their power-spectrum rows are 16 banks apart
            const int fl = 16 * half + MFCC_WARPS * it + warp;  // frame index inside the block
            const int t = t0 + fl;
            const bool valid = t < n_frames;
            const int fo = WW_HOP * t + origin_off;  // signal sample index of frame point n = 0
            const bool interior = valid && (fo + 95 >= 0) && (fo + 415 < L);
            // both frames of this warp lie past the end of the signal (short signals, streaming sessions that add one or
            // two frames per push): nothing to transform; their power rows stay stale and their outputs are never stored
            if (!__any_sync(0xffffffffu, valid)) continue;
            float* ps = pw + ROWS::p_row(fl);

            cpx v[16];
    #pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = cpk(0.f, 0.f);
            if (!EARLY_EDGE && block_has_edge && __any_sync(0xffffffffu, valid && !interior)) mbar_wait(&bars[3], edge_uses & 1);

            // this frame's staging half: frames 16*half.. live in half `half` (fl = 16*half + ...)
            const TIN* spcm = reinterpret_cast<const TIN*>(pcm_buf + half * SM::HALF_STRIDE);
            const int org = half ? org1 : org0;
            const float pre = a.preemph;
            if (interior) {
                // complex point m = 16*n1 + l16 (n1 = 3..12) <-> samples fo + 2m, fo + 2m + 1
                const int base = fo - org + 2 * l16;  // smem sample index of m = l16 (even: fo - org is a multiple of 8)
                const cpx mpre = cpk(-pre, -pre);
                const uint32_t* p32b = reinterpret_cast<const uint32_t*>(spcm) + (base >> 1);
    #pragma unroll
                for (int n1 = 3; n1 <= 12; ++n1) {
                    const cpx w = wreg[n1 - 3];
                    float x0, x1, xm1;
                    if constexpr (sizeof(TIN) == 2 && WW_IPRE) {
                        const uint32_t cur = p32b[16 * n1], prv = p32b[16 * n1 - 1];
                        // {x[2m-1], x[2m]} and {x[2m], x[2m+1]} as int16 pairs . {-97, 100}
                        const int y0 = __dp2a_lo((int)__byte_perm(prv, cur, 0x5432), 0x649f, 0);
                        const int y1 = __dp2a_lo((int)cur, 0x649f, 0);
                        v[n1] = p_mul(w, cpk(__int2float_rn(y0), __int2float_rn(y1)));   // w carries the 1/100
                        continue;
                    }
                    if constexpr (sizeof(TIN) == 2) {
                        const uint32_t cur = p32b[16 * n1], prv = p32b[16 * n1 - 1];
#if WW_I2FP
                        // sign-extend with PRMT / SHF and convert with I2FP.F32.S32 (FMA-side pipe) instead of the
                        // quarter-rate XU conversion I2F.S16
                        int i0, i1, im1;
                        asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(i0) : "r"(cur));
                        asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(i1) : "r"(cur));
                        asm("prmt.b32 %0, %1, 0, 0xbb32;" : "=r"(im1) : "r"(prv));
                        x0 = __int2float_rn(i0);
                        x1 = __int2float_rn(i1);
                        xm1 = __int2float_rn(im1);
#else
                        x0 = static_cast<float>(static_cast<int16_t>(cur & 0xffffu));
                        x1 = static_cast<float>(static_cast<int16_t>(cur >> 16));
                        xm1 = static_cast<float>(static_cast<int16_t>(prv >> 16));
#endif
                    } else {
                        const float* pf = reinterpret_cast<const float*>(spcm) + (base + 32 * n1);
                        const float2 c2 = *reinterpret_cast<const float2*>(pf);
                        x0 = c2.x;
                        x1 = c2.y;
                        xm1 = pf[-1];
                    }
                    // {w0 (x0 - pre x[-1]), w1 (x1 - pre x0)}
                    v[n1] = p_mul(w, p_fma(mpre, cpk(xm1, x0), cpk(x0, x1)));
                }
            } else if (valid) {
                // edge frame: taps were pre-emphasised into `edge` (slot 0: frame 0, slot 1: the tail frame)
                const float* ep = edge + ((EARLY_EDGE ? (int)(iter & 1) != 0 : t != 0) ? WW_WIN : 0) + 2 * l16;
    #pragma unroll
                for (int n1 = 3; n1 <= 12; ++n1) {
                    const cpx w = wreg[n1 - 3];
                    const cpx y = *reinterpret_cast<const cpx*>(ep + 32 * (n1 - 3));
                    v[n1] = p_mul(w, y);   // WW_IPRE: the edge taps were stored x 100 (emph_scale)
                }
            }

            // pass 1: DFT16 over n1, twiddle W256^(l16*k1), transpose through smem
            fft16<true>(v);
#if WW_TW1_HALF
    #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 tw = s_tw1[16 * j + l16];
                const cpx ta = cpk(tw.x, tw.y), tb = cpk(tw.z, tw.w);
                if (j > 0) v[2 * j] = p_cmulc(v[2 * j], ta);
                v[2 * j + 1] = p_cmulc(v[2 * j + 1], tb);
                v[2 * j + 8] = p_cmulc(v[2 * j + 8], j > 0 ? p_cmulc(tw1_8, ta) : tw1_8);
                v[2 * j + 9] = p_cmulc(v[2 * j + 9], p_cmulc(tw1_8, tb));
            }
#else
    #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 tw = j < WW_TW1_REGS ? tw1_r[j < WW_TW1_REGS ? j : 0] : s_tw1[16 * j + l16];
                if (j > 0) v[2 * j] = p_cmul(v[2 * j], tw.x, tw.y);
                v[2 * j + 1] = p_cmul(v[2 * j + 1], tw.z, tw.w);
            }
#endif
    #pragma unroll
            for (int k1 = 0; k1 < 16; ++k1)
                *reinterpret_cast<cpx*>(exch + k1 * EXCH_ROW_BYTES + l16 * 8) = v[k1];
            __syncwarp();
    #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(exch + l16 * EXCH_ROW_BYTES + j * 16);
                v[2 * j].v = q.x;
                v[2 * j + 1].v = q.y;
            }
            __syncwarp();
            // pass 2: DFT16 over n2 -> Z[l16 + 16*k2] = v[k2]
            fft16<false>(v);
            // P still holds the previous block until every warp has finished its mel stage
            if (mel_pending) {
                mbar_wait(&bars[2], mel_uses & 1);
                mel_pending = false;
            }
            // real-FFT split: pair (k, 256-k), k = l16 + 16*i -> 4*|X[k]|^2 and 4*|X[256-k]|^2.  Z[256-k] lives in lane
            // 16-l16, register 15-i (lane 0 pairs with itself: register 16-i, and Z[256] = Z[0])
            const int partner = (16 - l16) & 15;
            // lane that holds the partner row of the SAME frame
            const int plane = WW_LANE_INTERLEAVE ? 2 * partner + half : 16 * half + partner;
    #pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = l16 + 16 * i;
                const cpx za = v[i];
#if WW_PRESEL
                // lane 0 pairs with itself (register 16 - i, Z[256] = Z[0]): it offers that register to its own shuffle
                const float2 src = cunpk(l16 == 0 ? v[(16 - i) & 15] : v[15 - i]);
                const cpx zb = cpk(__shfl_sync(0xffffffffu, src.x, plane), __shfl_sync(0xffffffffu, src.y, plane));
#else
                const float2 src = cunpk(v[15 - i]);
                cpx zb = cpk(__shfl_sync(0xffffffffu, src.x, plane), __shfl_sync(0xffffffffu, src.y, plane));
                if (l16 == 0) zb = (i == 0) ? v[0] : v[(16 - i) & 15];
#endif
#if WW_TW2_COMPUTE
                // W512^(l16 + 16 i) = W512^l16 * W32^i
                constexpr float W32C[8] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
                constexpr float W32S[8] = {0.f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                                           -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
                const float2 w = i == 0 ? tw2_0 : cunpk(p_cmul(cpk(tw2_0.x, tw2_0.y), W32C[i], W32S[i]));
#else
                const float2 w = s_tw2[k];
#endif
                const cpx e = p_add(za, p_conj(zb));                 // (za.x + zb.x, za.y - zb.y)
                const cpx o = p_fma(cswap(za), pm1, cswap(zb));      // (za.y + zb.y, zb.x - za.x)
                const float2 ef = cunpk(e), tf = cunpk(p_cmul(o, w.x, w.y));
                // X[k] = e + tt, X[256-k] = conj(e - tt): real parts {e.x + tt.x, e.x - tt.x} and imaginary parts
                // {e.y + tt.y, e.y - tt.y} as packed pairs (scalar operands are broadcast by the instruction), so the
                // two power bins come out of one FMUL2 + one FFMA2
                const cpx xr = p_fma(cpk(tf.x, tf.x), pm1, cpk(ef.x, ef.x));
                const cpx xi = p_fma(cpk(tf.y, tf.y), pm1, cpk(ef.y, ef.y));
                const float2 pp = cunpk(p_fma(xi, xi, p_mul(xr, xr)));
                ps[k] = pp.x;
                ps[256 - k] = pp.y;
            }
            if (l16 == 0) {
                const float2 z8 = cunpk(v[8]);
                ps[128] = 4.f * fmaf(z8.x, z8.x, z8.y * z8.y);
            }
            __syncwarp();
        }
        if (prev_valid) ++mel_uses;
        if (block_has_edge && !EARLY_EDGE) ++edge_uses;
        // one CTA barrier per block: P(k) is complete; LM(k-1), the edge taps and PCM(k) are free again
        if constexpr (EARLY_EDGE) {
            __syncwarp();
#ifdef WW_MFCC_STATS
            const long long t_arr = clock64();
#endif
            if (lane == 0) mbar_arrive(&bars[4]);
            // while the slower warps finish: the edge taps of the next block (its PCM was staged an iteration ago) go to
            // the tap slot this block did not use
            if (blk_id + stride < a.n_blocks) {
                const int nb = (int)((iter + 1) & 1);
                mbar_wait(&bars[nb], (uint32_t)(((iter + 1) >> 1) & 1));
                fill_edge_ahead(bi_nxt, smem + SM::OFF_PCM + nb * SM::PCM_BYTES, nb);
            }
            mbar_wait(&bars[4], (uint32_t)(iter & 1));
#ifdef WW_MFCC_STATS
            if (lane == 0 && !PIPE::FUSED) g_mfcc_slack[(blockIdx.x & 1023) * 8 + warp] += (unsigned long long)(clock64() - t_arr);
#endif
        } else {
            pipe.sync();
        }
        pub_valid = prev_valid;
        if constexpr (PIPE::FUSED) v_sig += stride_q;
        prev_valid = have;
        prev_sig = sig;
        prev_t0 = t0;
        sig_cur = sig_nxt;
        bi_cur = bi_nxt;
    }
    if constexpr (PIPE::FUSED) {
        if (pub_valid && tid == MFCC_THREADS - 32) pipe.publish(v_sig - 2 * stride_q);
    }
}

template <typename TIN, int MEL, bool CLIP = false>
__global__ void __launch_bounds__(MFCC_THREADS, 2) mfcc_kernel(const __grid_constant__ MfccArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_launch_dependents();
    MfccSolo pipe;
#if WW_PCM_EVICT_FIRST
    pipe.pol_ = l2_policy_evict_first();
#endif
    mfcc_body<TIN, MEL, CLIP, MfccSolo>(a, smem, pipe);
}

}  // namespace ww
