// int8 power-of-two twin of LightweightKWS on the tensor cores: tcgen05.mma kind::i8 (SASS UTCIMMA), int8 operands
// from shared memory, int32 accumulators in TMEM.  Integer-exact: same results, bit for bit, as cnn_i8_kernel
// (ww_cnn_i8.cuh) and as the reference's shipped known-answer vector (ml_models/xiaoa.info:3153-3224).
//
// Reference: ml_models/xiaoa.info (int8 weights :31-3136, exponents :3139-3150), ml_models/xiaoa.json:5-20, device
// requantisation path main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:128-131,200-220.
//
// Structure = cnn_tc_kernel (ww_cnn_tc.cuh): one persistent CTA per SM, FOUR independent 4-warp groups, each
// scoring 8 windows per iteration through conv1/conv2 (even/odd de-interleaved operands, per-thread MaxPool),
// conv3 (channels on the TMEM lanes) and fc1.  Differences that int8 brings:
//   * a 16-byte K chunk holds 16 channels and one MMA spans K = 32: conv1 (13 -> 16 channels) multiplies a second,
//     arbitrary chunk by zero weights; conv2 is one MMA per tap, conv3 two, fc1 four (28 MMAs per octet, not 44)
//   * requantisation (ReLU, right shift with round-half-to-even, int8 saturation) happens in the TMEM epilogues;
//     it is monotonic, so MaxPool runs first on the int32 accumulators and one requantisation serves both taps
//   * activations are a quarter of the fp16 kernel's bytes: 27 KB per group, weights 49 KB -> four groups fit
#pragma once
#include "ww_cnn_i8.cuh"
#include "ww_cnn_tc.cuh"

namespace ww {

constexpr int I8T_GROUPS = 4;
constexpr int I8T_THREADS = 128 * I8T_GROUPS;
constexpr int I8T_CLIPS = 8;

constexpr int I8_A1P_ROWS = 32 * I8T_CLIPS + 2, I8_A2P_ROWS = 16 * I8T_CLIPS + 2, I8_X3_ROWS = 16 * I8T_CLIPS + 2;
constexpr int I8_A1_PAR = I8_A1P_ROWS * 16;                      // one 16-channel chunk per parity tile
constexpr int I8_A2_LBO = I8_A2P_ROWS * 16, I8_A2_PAR = 2 * I8_A2_LBO;
constexpr int I8_X3_LBO = I8_X3_ROWS * 16;
constexpr int I8_G_LBO = 16 * 16;
constexpr int I8_ACT_A1 = 0;
constexpr int I8_ACT_A2 = I8_ACT_A1 + 2 * I8_A1_PAR;
constexpr int I8_ACT_X3 = I8_ACT_A2 + 2 * I8_A2_PAR;
constexpr int I8_ACT_G = I8_ACT_X3 + 4 * I8_X3_LBO;
constexpr int I8_ACT_BYTES = I8_ACT_G + 8 * I8_G_LBO;            // 26 944
static_assert(2 * I8_A2_PAR >= I8_A1_PAR, "the throw-away second K chunk of conv1's odd tile must stay inside the group");

// weight blob: K-major 16-byte chunks, element (row n, k) at (k/16)*rows*16 + n*16 + k%16
constexpr int I8_W1_LBO = 32 * 16, I8_W1_TAP = 2 * I8_W1_LBO;
constexpr int I8_W2_LBO = 64 * 16, I8_W2_TAP = 2 * I8_W2_LBO;
constexpr int I8_W3_LBO = 128 * 16, I8_W3_TAP = 4 * I8_W3_LBO;
constexpr int I8_WF1_LBO = 128 * 16;
constexpr int I8_W1 = 0;
constexpr int I8_W2 = I8_W1 + 3 * I8_W1_TAP;
constexpr int I8_W3 = I8_W2 + 3 * I8_W2_TAP;
constexpr int I8_WF1 = I8_W3 + 3 * I8_W3_TAP;
constexpr int I8_W_BYTES = I8_WF1 + 8 * I8_WF1_LBO;              // 50 176

constexpr int I8T_OFF_BAR = 0;                                   // mbarrier[4] + tmem base at +32
constexpr int I8T_OFF_PART = 64;                                 // fc2 partial sums [group][2][8][8] ints
constexpr int I8T_OFF_FC2 = I8T_OFF_PART + I8T_GROUPS * 2 * 8 * 8 * 4;
constexpr int I8T_OFF_W = I8T_OFF_FC2 + TC_MAX_CLASSES * 64 * 4;
constexpr int I8T_OFF_ACT = I8T_OFF_W + I8_W_BYTES;
constexpr int I8T_SMEM = I8T_OFF_ACT + I8T_GROUPS * I8_ACT_BYTES;
static_assert(I8T_OFF_W % 16 == 0 && I8T_OFF_ACT % 16 == 0 && I8_ACT_BYTES % 16 == 0, "UMMA operand alignment");
static_assert(I8T_SMEM <= 232448, "shared memory budget");
constexpr int I8T_GROUP_COLS = 128;   // conv accumulators [0,128); fc1 reuses [0,16) after the conv3 epilogue
constexpr int I8T_TMEM_COLS = 512;

struct I8TcArgs {
    const signed char* x;    // [n][13][63] int8 at the model-input exponent (coef-major), or null when `feats` is given
    // float-feature input (the whole device path in one launch): feats[win*win_stride + coef*coef_stride +
    // frame*frame_stride] -> int8 rounding + device CMVN (tc_cmvn_device) -> model input at exponent -4
    const float* feats;
    long long win_stride, coef_stride, frame_stride, group_windows, group_stride;
    long long n_windows;
    signed char* out;        // [n][C] int8 at the output exponent (may be null)
    float* logits_f;         // [n][C] dequantised logits out * 2^exp_out (may be null)
    unsigned char* decisions;  // class 0, may be null
    int decide_mode;
    float threshold;
    float out_scale;         // 2^exp_out
    const uint4* wblob;      // I8_W_BYTES
    const signed char* fc2;  // [C][64]
    int num_classes;
    int sh1, sh2, sh3, shf1, shf2, gap_num_shift;
};

__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    // c_format S32 (bits 4-5 = 2), a/b format signed 8-bit (1 at bits 7 and 10), both K-major, N>>3 at 17, M>>4 at 24
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ReLU'd accumulator (>= 0) -> int8: right shift with round-half-to-even, saturate at 127.
// (acc + 2^(s-1) - 1 + ((acc >> s) & 1)) >> s  ==  requant_i8(acc, s) for s > 0 (ww_cnn_i8.cuh)
__device__ __forceinline__ int rq_pos(int acc, int shift, int hm1) {
    if (shift <= 0) return min(127, acc << (-shift));
    return min(127, (acc + hm1 + ((acc >> shift) & 1)) >> shift);
}
// four int32 -> four int8 with saturation, two I2IP.S8.S32.SAT: cvt.pack puts sat(a) in byte 1, sat(b) in byte 0 and
// the low half of c above them
__device__ __forceinline__ uint32_t pack_b4(int x0, int x1, int x2, int x3) {
    uint32_t t, d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(x3), "r"(x2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x1), "r"(x0), "r"(t));
    return d;
}
// rq_pos without the clamp: the saturating pack that follows applies it
__device__ __forceinline__ int rq_pos_nosat(int acc, int shift, int hm1) {
    if (shift <= 0) return acc << (-shift);
    return (acc + hm1 + ((acc >> shift) & 1)) >> shift;
}
__device__ __forceinline__ uint4 pack_b16(const int* v) {
    return make_uint4(pack_b4(v[0], v[1], v[2], v[3]), pack_b4(v[4], v[5], v[6], v[7]), pack_b4(v[8], v[9], v[10], v[11]),
                      pack_b4(v[12], v[13], v[14], v[15]));
}
__device__ __forceinline__ void tmem_ld32x2_i(uint32_t ta, uint32_t tb, int (&va)[32], int (&vb)[32]) {
    float fa[32], fb[32];
    tmem_ld32x2(ta, tb, fa, fb);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        va[i] = __float_as_int(fa[i]);
        vb[i] = __float_as_int(fb[i]);
    }
}

__global__ void __launch_bounds__(I8T_THREADS, 1) cnn_i8_tc_kernel(const __grid_constant__ I8TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + I8T_OFF_BAR + 32);
    int* sfc2 = reinterpret_cast<int*>(smem + I8T_OFF_FC2);
    unsigned char* sW = smem + I8T_OFF_W;

    // warp index through a shuffle: known warp-uniform to the compiler, so the UMMA descriptors are built on the uniform
    // datapath (see cnn_tc_body)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = warp_index_uniform(tid);
    const int group = warp >> 2, q4 = warp & 3, tig = tid & 127;
    const int C = a.num_classes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + I8T_OFF_BAR) + group;
    int* part = reinterpret_cast<int*>(smem + I8T_OFF_PART) + group * (2 * 8 * 8);
    unsigned char* act = smem + I8T_OFF_ACT + group * I8_ACT_BYTES;
    unsigned char* sA1 = act + I8_ACT_A1;
    unsigned char* sA2 = act + I8_ACT_A2;
    unsigned char* sX3 = act + I8_ACT_X3;
    unsigned char* sG = act + I8_ACT_G;

    for (int i = tid; i < (I8T_SMEM - I8T_OFF_ACT) / 16; i += I8T_THREADS)
        reinterpret_cast<uint4*>(smem + I8T_OFF_ACT)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < I8_W_BYTES / 16; i += I8T_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(a.wblob + i);
    for (int i = tid; i < C * 64; i += I8T_THREADS) sfc2[i] = a.fc2[i];
    if (tid == 0) {
        for (int g = 0; g < I8T_GROUPS; ++g) mbar_init(reinterpret_cast<uint64_t*>(smem + I8T_OFF_BAR) + g, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)I8T_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0) + (uint32_t)(group * I8T_GROUP_COLS);
    const uint32_t sA1a = smem_u32(sA1), sA2a = smem_u32(sA2), sX3a = smem_u32(sX3), sGa = smem_u32(sG), sWa = smem_u32(sW);
    const uint32_t tlane = (uint32_t)(32 * q4) << 16;
    uint32_t phase = 0;
    const int sh1 = a.sh1, sh2 = a.sh2, sh3 = a.sh3, shf1 = a.shf1;
    const int hm1_1 = sh1 > 0 ? (1 << (sh1 - 1)) - 1 : 0, hm1_2 = sh2 > 0 ? (1 << (sh2 - 1)) - 1 : 0;
    const int hm1_3 = sh3 > 0 ? (1 << (sh3 - 1)) - 1 : 0, hm1_f = shf1 > 0 ? (1 << (shf1 - 1)) - 1 : 0;

    const long long n_oct = (a.n_windows + I8T_CLIPS - 1) / I8T_CLIPS;
    const long long oct_stride = (long long)gridDim.x * I8T_GROUPS;
    // as in cnn_tc_kernel: the next octet of a flat [n][13][63] float batch is pulled into L2 while this one is computed
    const bool flat = a.feats && a.group_windows == 0 && a.frame_stride == 1 && a.coef_stride == WW_WINDOW_FRAMES &&
                      a.win_stride == WW_N_MFCC * WW_WINDOW_FRAMES && (reinterpret_cast<uintptr_t>(a.feats) & 15) == 0;
#pragma unroll 1
    for (long long oct = (long long)blockIdx.x * I8T_GROUPS + group; oct < n_oct; oct += oct_stride) {
        if (WW_TC_PREFETCH && flat && tig == 0 && oct + oct_stride < n_oct) {
            long long wins = a.n_windows - (oct + oct_stride) * I8T_CLIPS;
            wins = wins < I8T_CLIPS ? wins : I8T_CLIPS;
            const uint32_t bytes = (uint32_t)(wins * WW_N_MFCC * WW_WINDOW_FRAMES * 4) & ~15u;
            if (bytes) bulk_prefetch_l2(a.feats + (oct + oct_stride) * (long long)(I8T_CLIPS * WW_N_MFCC * WW_WINDOW_FRAMES), bytes);
        }
        // ================= S0: model input -> A1 rows (lane <-> frame), two windows per warp =================
        if (a.feats) {
            // float features: int8 rounding + device-style CMVN here, so the device path is one launch
            TcWin w2[2];
            if (flat) {   // immediate-offset loads, as in cnn_tc_body
                const long long w0 = oct * I8T_CLIPS + 2 * q4;
                const float* wbase = a.feats + w0 * (long long)(WW_N_MFCC * WW_WINDOW_FRAMES);
                tc_load_window_flat(wbase, w0 < a.n_windows, lane, w2[0]);
                tc_load_window_flat(wbase + WW_N_MFCC * WW_WINDOW_FRAMES, w0 + 1 < a.n_windows, lane, w2[1]);
            } else {
                tc_load_window(a, oct * I8T_CLIPS + 2 * q4, lane, w2[0]);       // both windows' loads in flight
                tc_load_window(a, oct * I8T_CLIPS + 2 * q4 + 1, lane, w2[1]);
            }
#pragma unroll
            for (int ww_ = 0; ww_ < 2; ++ww_) {
                const int slot = 2 * q4 + ww_;
                TcWin& w = w2[ww_];
                tc_cmvn_device(w, lane);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int t = lane + 32 * hf;
                    if (t >= WW_WINDOW_FRAMES) continue;
                    int v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = q < WW_N_MFCC ? (int)((hf ? w.x1[q] : w.x0[q]) * 16.f) : 0;
                    *reinterpret_cast<uint4*>(sA1 + (t & 1) * I8_A1_PAR + (1 + 32 * slot + (t >> 1)) * 16) = pack_b16(v);
                }
            }
        } else {
#pragma unroll
            for (int ww_ = 0; ww_ < 2; ++ww_) {
                const int slot = 2 * q4 + ww_;
                const long long win = oct * I8T_CLIPS + slot;
                const bool live = win < a.n_windows;
                const signed char* src = a.x + (live ? win : 0) * (WW_N_MFCC * WW_WINDOW_FRAMES);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int t = lane + 32 * hf;
                    if (t >= WW_WINDOW_FRAMES) continue;   // frame 63 does not exist: its (odd-tile) row stays zero
                    int v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = (q < WW_N_MFCC && live) ? (int)src[q * WW_WINDOW_FRAMES + t] : 0;
                    *reinterpret_cast<uint4*>(sA1 + (t & 1) * I8_A1_PAR + (1 + 32 * slot + (t >> 1)) * 16) = pack_b16(v);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= conv1: 2 row tiles x {even, odd} x 3 taps, K = 32 (channels 16..31: zero weights) ======
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_i8(128, 32);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int par = 0; par < 2; ++par)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int src_par = par ? (r == 1) : (r != 1);
                        const int shift = par ? (r == 2) : -(r == 0);
                        umma_i8(tmem + 64 * i + 32 * par,
                                umma_desc_kmajor(sA1a + src_par * I8_A1_PAR + (1 + 128 * i + shift) * 16, I8_A1_PAR),
                                umma_desc_kmajor(sWa + I8_W1 + r * I8_W1_TAP, I8_W1_LBO), idesc, r > 0);
                    }
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
            int ve[32], vo[32];
            tmem_ld32x2_i(tmem + tlane + 64 * i, tmem + tlane + 64 * i + 32, ve, vo);
            const int g = 128 * i + 32 * q4 + lane;
            const int w = g >> 5, j = g & 31;
            int mine[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) mine[c] = rq_pos_nosat(max(max(ve[c], vo[c]), 0), sh1, hm1_1);
            if (j < 31) {
                unsigned char* dst = sA2 + (j & 1) * I8_A2_PAR + (1 + 16 * w + (j >> 1)) * 16;
                *reinterpret_cast<uint4*>(dst) = pack_b16(mine);
                *reinterpret_cast<uint4*>(dst + I8_A2_LBO) = pack_b16(mine + 16);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= conv2: {even, odd} x 3 taps, K = 32 =================
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_i8(128, 64);
#pragma unroll
            for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int src_par = par ? (r == 1) : (r != 1);
                    const int shift = par ? (r == 2) : -(r == 0);
                    umma_i8(tmem + 64 * par, umma_desc_kmajor(sA2a + src_par * I8_A2_PAR + (1 + shift) * 16, I8_A2_LBO),
                            umma_desc_kmajor(sWa + I8_W2 + r * I8_W2_TAP, I8_W2_LBO), idesc, r > 0);
                }
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        {
            const int g = 32 * q4 + lane;
            const bool valid = (g & 15) < 15;
            unsigned char* dst = sX3 + (1 + g) * 16;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                int ve[32], vo[32];
                tmem_ld32x2_i(tmem + tlane + 32 * hh, tmem + tlane + 64 + 32 * hh, ve, vo);
                int mine[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) mine[c] = valid ? rq_pos_nosat(max(max(ve[c], vo[c]), 0), sh2, hm1_2) : 0;
                *reinterpret_cast<uint4*>(dst + (2 * hh) * I8_X3_LBO) = pack_b16(mine);
                *reinterpret_cast<uint4*>(dst + (2 * hh + 1) * I8_X3_LBO) = pack_b16(mine + 16);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= conv3 (channels on M): 3 taps x 2 K-steps, N = 128 positions =================
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_i8(128, 128);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma_i8(tmem, umma_desc_kmajor(sWa + I8_W3 + r * I8_W3_TAP + ks * 2 * I8_W3_LBO, I8_W3_LBO),
                            umma_desc_kmajor(sX3a + r * 16 + ks * 2 * I8_X3_LBO, I8_X3_LBO), idesc, (r | ks) > 0);
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        // thread = channel o: per window ReLU + requantise + MaxPool + exact mean of the 7 pooled steps -> G (int8)
        {
            const int o = 32 * q4 + lane;
            const int gsh = a.gap_num_shift;
            const int den = gsh >= 0 ? 7 : (7 << (-gsh));
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                int va[32], vb[32];
                tmem_ld32x2_i(tmem + tlane + 64 * h, tmem + tlane + 64 * h + 32, va, vb);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int* v = cc < 2 ? va + 16 * cc : vb + 16 * (cc - 2);
                    int s = 0;
#pragma unroll
                    for (int j = 0; j < 7; ++j) s += rq_pos(max(max(v[2 * j], v[2 * j + 1]), 0), sh3, hm1_3);
                    // nearest integer (ties to even) of s * 2^gsh / 7, s >= 0   (cnn_i8_kernel's global average)
                    const int num = gsh >= 0 ? (s << gsh) : s;
                    const int n2 = 2 * num, d2 = 2 * den;
                    const int fl = n2 / d2, rem = n2 - fl * d2;
                    const int qv = min(127, fl + ((rem > den || (rem == den && (fl & 1))) ? 1 : 0));
                    const int clip = 4 * h + cc;
                    sG[(o >> 4) * I8_G_LBO + clip * 16 + (o & 15)] = (unsigned char)qv;
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= fc1: [128(64) x 128] . [16 x 128]^T, 4 K-steps, accumulators in columns [0,16) ==========
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_i8(128, 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                umma_i8(tmem, umma_desc_kmajor(sWa + I8_WF1 + ks * 2 * I8_WF1_LBO, I8_WF1_LBO),
                        umma_desc_kmajor(sGa + ks * 2 * I8_G_LBO, I8_G_LBO), idesc, ks > 0);
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        if (q4 < 2) {
            float hf[16];
            tmem_ld16(tmem + tlane, hf);
            const int o = 32 * q4 + lane;
            int h[8];
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) h[c8] = rq_pos(max(__float_as_int(hf[c8]), 0), shf1, hm1_f);
            for (int c = 0; c < C; ++c) {
                const int w = sfc2[c * 64 + o];
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int s = __reduce_add_sync(0xffffffffu, h[c8] * w);
                    if (lane == 0) part[(q4 * 8 + c) * 8 + c8] = s;
                }
            }
        }
        tc_fence_before();
        group_sync(group);
        if (tig < 8 * C) {
            const int c8 = tig & 7, c = tig >> 3;
            const long long win = oct * I8T_CLIPS + c8;
            if (win < a.n_windows) {
                const int oq = requant_i8(part[(0 * 8 + c) * 8 + c8] + part[(1 * 8 + c) * 8 + c8], a.shf2);
                if (a.out) a.out[win * C + c] = (signed char)oq;
                const float s = (float)oq * a.out_scale;   // exact: int8 times a power of two
                if (a.logits_f) a.logits_f[win * C + c] = s;
                if (c == 0 && a.decisions) {
                    unsigned char d = 0;
                    if (a.decide_mode == DECIDE_LOGIT) d = s > a.threshold;
                    else if (a.decide_mode == DECIDE_DEVICE) d = (1.f / (1.f + expf(-s)) * 100.f) >= a.threshold;
                    a.decisions[win] = d;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)I8T_TMEM_COLS)
                     : "memory");
    }
}

// ---- host: int8 weight blob in UMMA K-major layout (16 int8 per chunk) ---------------------------------
// conv weights arrive as the [cin][tap][cout] int8 arrays of cnn_i8_kernel, fc1 as [64][128]
inline void i8tc_build_blob(std::vector<unsigned char>& blob, const signed char* w1t, const signed char* w2t,
                            const signed char* w3t, const signed char* fc1) {
    blob.assign(I8_W_BYTES, 0);
    auto put = [&](int base, int rows, int n, int k, signed char v) {
        blob[(size_t)base + (size_t)(k / 16) * rows * 16 + (size_t)n * 16 + (k % 16)] = (unsigned char)v;
    };
    for (int r = 0; r < 3; ++r) {
        for (int i = 0; i < 13; ++i)
            for (int o = 0; o < 32; ++o) put(I8_W1 + r * I8_W1_TAP, 32, o, i, w1t[(i * 3 + r) * 32 + o]);
        for (int i = 0; i < 32; ++i)
            for (int o = 0; o < 64; ++o) put(I8_W2 + r * I8_W2_TAP, 64, o, i, w2t[(i * 3 + r) * 64 + o]);
        for (int i = 0; i < 64; ++i)
            for (int o = 0; o < 128; ++o) put(I8_W3 + r * I8_W3_TAP, 128, o, i, w3t[(i * 3 + r) * 128 + o]);
    }
    for (int o = 0; o < 64; ++o)
        for (int i = 0; i < 128; ++i) put(I8_WF1, 128, o, i, fc1[o * 128 + i]);
}

}  // namespace ww
