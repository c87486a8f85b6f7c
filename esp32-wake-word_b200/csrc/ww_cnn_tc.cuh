// CMVN + LightweightKWS forward on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Same contract as cnn_fp32_kernel (ww_cnn.cuh): replaces normalize_mfcc (extract_mfcc.py:47-88) /
// device CMVN (esp_wake_word_detector.cpp:179-211), LightweightKWS.forward (wakeModel.py:29-34) and the
// decision (ml_models/main.py:53, esp_wake_word_detector.cpp:226-245) for 63-frame windows.
//
// One persistent CTA per SM runs TC_GROUPS (4) independent 4-warp groups; each group scores EIGHT windows per iteration, so
// one group's TMEM epilogues (CUDA cores) overlap the other group's MMAs and waits.  Every layer is an implicit GEMM
// issued by a single thread of the group with tcgen05.mma (kind::f16: fp16 operands, fp32 accumulation in TMEM):
//   conv1  D[512 pos x 32]  = sum_tap A1[pos+tap][16] . W1_tap[32][16]^T     4 tiles of M=128, K=16 per tap
//   conv2  D[256 pos x 64]  = sum_tap A2[pos+tap][32] . W2_tap[64][32]^T     2 tiles of M=128
//          (even and odd output positions are separate tiles, see A1P_ROWS below)
//   conv3  D[128 ch x 128 pos] = sum_tap W3_tap[128][64] . X3[pos+tap][64]^T   (roles swapped: channel = TMEM lane,
//                                                            so MaxPool and the global average are per-thread)
//   fc1    D[128(64) x 16(8 windows)] = WF1[128][128] . G[16][128]^T
// Operands live in shared memory in the canonical K-major, no-swizzle UMMA layout with an 8-row group stride of
// 128 B, i.e. row r of a 16-byte K-chunk sits at chunk_base + 16*r.  That makes the k=3 convolution im2col-free:
// tap r is the same tile with its start address advanced by r rows (16*r bytes); a zero row between consecutive
// windows provides the padding.  ReLU + MaxPool(2) are fused into the TMEM->register epilogue, which writes the
// next layer's operand directly in that layout (fp16).  fc2 (64 -> C) is a warp reduction on the fc1 epilogue.
// Windows whose logit lands within `band` of the decision threshold are appended to a re-score list that the
// fp32 kernel then recomputes exactly: decisions are those of the fp32 path.
#pragma once
#include <cuda_fp16.h>

#include "ww_cnn.cuh"
#include "ww_common.cuh"

namespace ww {

constexpr int TC_GROUPS = 4;      // independent 4-warp groups per CTA
constexpr int TC_THREADS = 128 * TC_GROUPS;
constexpr int TC_CLIPS = 8;
constexpr int TC_MAX_CLASSES = 8;

// conv1 / conv2 operands are stored DE-INTERLEAVED: one tile holds the even positions of every window, a second
// tile the odd ones (row 1 + 32*w + j <-> position 2j / 2j+1 of window w; row 0 and the last row of each odd
// block stay zero = conv padding).  A k=3 tap is then still a row-shifted view of one of the two tiles, the conv
// outputs at even and odd positions come out of separate MMAs into separate TMEM columns of the SAME lane, and
// MaxPool(2) is a per-thread max -- no shuffles in the epilogues.
constexpr int A1P_ROWS = 32 * TC_CLIPS + 2;  // 258 rows per parity tile (63 frames -> 32 even + 31 odd positions)
constexpr int A2P_ROWS = 16 * TC_CLIPS + 2;  // 130 (31 positions -> 16 even + 15 odd)
constexpr int X3_ROWS = 16 * TC_CLIPS + 2;   // 130: natural order (conv3 pools along TMEM columns)
constexpr int G_ROWS = 16;                   // fc1 B operand: 8 windows + 8 zero rows (N must be a multiple of 16)
// Bank staggers (ncu: 31 % of the kernel's shared-memory wavefronts were bank-conflict replays).  The conv1 epilogue
// writes one 16-byte row per lane, even lanes into the even-position tile of A2 and odd lanes into the odd one: with
// A2_PAR a multiple of 128 both halves of a quarter-warp hit the same banks (2-way), so the odd tile starts 64 bytes
// later (A1_PAR = 8256 already is 64 mod 128).  The conv3 epilogue writes G two bytes per lane, eight lanes per
// 16-byte row, the next eight lanes one K chunk further: a chunk stride of 256 puts all four on the same banks
// (4-way), 272 spreads them.  UMMA only needs 16-byte multiples for tile bases and the K-chunk stride (LBO).
#ifndef WW_TC_STAGGER
#define WW_TC_STAGGER 1
#endif
constexpr int A1_LBO = A1P_ROWS * 16, A2_LBO = A2P_ROWS * 16, X3_LBO = X3_ROWS * 16, G_LBO = G_ROWS * 16 + (WW_TC_STAGGER ? 16 : 0);
constexpr int A1_PAR = 2 * A1_LBO, A2_PAR = 4 * A2_LBO + (WW_TC_STAGGER ? 64 : 0);  // bytes per parity tile
// fc1's A operand stores only its 64 real rows per K chunk: rows 64..127 of the M = 128 tile read the next chunk
// (finite weights; one zero chunk follows the last) and produce accumulator rows nobody reads
constexpr int W1_LBO = 32 * 16, W2_LBO = 64 * 16, W3_LBO = 128 * 16, WF1_LBO = 64 * 16;
constexpr int W1_TAP = 2 * W1_LBO, W2_TAP = 4 * W2_LBO, W3_TAP = 8 * W3_LBO;

// shared memory map (bytes).  The CTA runs TWO independent 4-warp groups, each scoring its own octet of windows
// with its own activation tiles, mbarrier and TMEM columns; the weights are shared.
constexpr int TC_GROUP_THREADS = TC_THREADS / TC_GROUPS;       // 128: one warp per TMEM lane quadrant
constexpr int TC_OFF_BAR = 0;                                  // mbarrier[TC_GROUPS] + tmem base (4) at +32
constexpr int TC_OFF_NORM = 64;                                // ||x||_F of the windows in flight [group][octet parity][8] floats
constexpr int TC_OFF_PART = TC_OFF_NORM + TC_GROUPS * 2 * 8 * 4;  // fc2 partial sums [group][2][8][8] floats
constexpr int TC_OFF_FC2 = TC_OFF_PART + TC_GROUPS * 2 * 8 * 8 * 4;  // fc2 weights [8][64] floats
constexpr int TC_OFF_W = TC_OFF_FC2 + TC_MAX_CLASSES * 64 * 4;  // weight blob (same layout as the device blob)
constexpr int TC_W1 = 0;
constexpr int TC_W2 = TC_W1 + 3 * W1_TAP;
constexpr int TC_W3 = TC_W2 + 3 * W2_TAP;
constexpr int TC_WF1 = TC_W3 + 3 * W3_TAP;
constexpr int TC_W_BYTES = TC_WF1 + 17 * WF1_LBO;              // 81 920 (16 chunks + one zero chunk)
constexpr int TC_OFF_ACT = TC_OFF_W + TC_W_BYTES;              // per-group activation tiles
// A1 (conv1 operand) and X3 (conv3 operand) share storage: A1 is dead once conv1 has completed, X3 is written by
// the conv2 epilogue.  The zero rows each of them relies on are re-written every octet (see S0 / epilogue 2).
// G (fc1 operand, written by the conv3 epilogue) lies over the start of A2's even-position tile, which is dead
// after conv2 and fully rewritten by the next conv1 epilogue; G's rows 8..15 (the padding of N = 16) then hold
// stale activations that only reach accumulator columns nobody reads.
constexpr int TC_ACT_A1 = 0;
constexpr int TC_ACT_X3 = 0;
constexpr int TC_ACT_A2 = (2 * A1_PAR > 8 * X3_LBO ? 2 * A1_PAR : 8 * X3_LBO);
constexpr int TC_ACT_G = TC_ACT_A2;
constexpr int TC_ACT_BYTES = TC_ACT_A2 + 2 * A2_PAR;           // 33 408
static_assert(16 * G_LBO <= A2_PAR, "G must stay inside A2's even-position tile");
constexpr int TC_SMEM = TC_OFF_ACT + TC_GROUPS * TC_ACT_BYTES;
static_assert(TC_OFF_W % 16 == 0 && TC_OFF_ACT % 16 == 0 && TC_ACT_A2 % 16 == 0 && TC_ACT_X3 % 16 == 0 &&
                  TC_ACT_G % 16 == 0 && TC_ACT_BYTES % 16 == 0,
              "UMMA operands need 16-byte alignment");
static_assert(TC_SMEM <= 232448, "shared memory budget");
constexpr int TC_GROUP_COLS = 128;  // per group: conv accumulators use columns [0,128); fc1 reuses [0,16) after epilogue 3
constexpr int TC_TMEM_COLS = 512;
static_assert(TC_GROUPS * TC_GROUP_COLS <= TC_TMEM_COLS, "TMEM columns");

struct TcArgs {
    const float* feats;  // feats[win*win_stride + coef*coef_stride + frame*frame_stride]
    long long win_stride, coef_stride, frame_stride;
    long long n_windows;
    long long group_windows, group_stride;  // as in CnnArgs
    int cmvn_mode, decide_mode;
    float threshold;        // as in CnnArgs
    // Guard band of the fp16-operand path (ww_api.cu:tc_calibrate).  The network is bias-free, so its rounding error is
    // proportional to the norm of the window it is fed: a window whose class-0 logit lies within
    //   bw = band + band_rel * ||x||_F      (band_rel != 0 only for CMVN_NONE; the CMVN modes fix ||z||_F)
    // of thr0 or thr1 (DECIDE_NONE: 0 and ln 4, the two decision rules of the reference; else the rule asked for, twice),
    // whose two largest class logits are closer than 2 bw, or whose norm exceeds norm_limit (fp16 range) is re-scored.
    float thr0, thr1;
    float band, band_rel, norm_limit;
    float* logits;          // [n][C]
    unsigned char* decisions;
    long long* rescore_list;  // may be null
    int* rescore_count;
    // Compact hand-over to the exact kernel (flat [n][13][63] batches and the fused clip kernel): a listed window's
    // features are copied to rescore_feat[position in the list] and the list entry is win + rescore_base, so that one
    // re-score launch can serve many CNN launches whose feature buffer has been recycled in between.  null: the list
    // indexes `feats` itself and the re-score follows this launch.
    float* rescore_feat;
    long long rescore_base;
    const uint4* wblob;     // TC_W_BYTES
    const float* fc2;       // [C][64]
    int num_classes;
    float* dbg;             // optional stage dump of the first octet (tests)
};

// ---- tcgen05 primitives -----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes) {
    // K-major, SWIZZLE_NONE: 8-row core matrices of 8 x 16 B; SBO = 128 B (rows 16 B apart), LBO = K-chunk stride
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    // c_format F32 (bits 4-5 = 1), a/b format F16 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// two 32-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t tb, float (&va)[32], float (&vb)[32]) {
    uint32_t r[32], q[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
          "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
          "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(tb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        va[i] = __uint_as_float(r[i]);
        vb[i] = __uint_as_float(q[i]);
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack_h8(const float* v) {
    return make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
}

// Sum each of 16 per-lane values over the 32 lanes with 16 shuffles (recursive halving): afterwards lane l holds
// the warp total of value index (l >> 1) & 15.
__device__ __forceinline__ float reduce16(const float (&v)[16], int lane) {
    float a8[8], a4[4], a2[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
            a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? a8[i] : a8[i + 4], keep = up ? a8[i + 4] : a8[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = up ? a4[i] : a4[i + 2], keep = up ? a4[i + 2] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool up = lane & 2;
    const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0];
    float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    return a1;
}

// reduce16 of two value sets in lock-step (the same operations per set, interleaved level by level)
__device__ __forceinline__ void reduce16x2(const float (&va)[16], const float (&vb)[16], int lane, float& ra, float& rb) {
    float a8[8], b8[8], a4[4], b4[4], a2[2], b2[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float sa = up ? va[i] : va[i + 8], ka = up ? va[i + 8] : va[i];
            const float sb = up ? vb[i] : vb[i + 8], kb = up ? vb[i + 8] : vb[i];
            a8[i] = ka + __shfl_xor_sync(0xffffffffu, sa, 16);
            b8[i] = kb + __shfl_xor_sync(0xffffffffu, sb, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sa = up ? a8[i] : a8[i + 4], ka = up ? a8[i + 4] : a8[i];
            const float sb = up ? b8[i] : b8[i + 4], kb = up ? b8[i + 4] : b8[i];
            a4[i] = ka + __shfl_xor_sync(0xffffffffu, sa, 8);
            b4[i] = kb + __shfl_xor_sync(0xffffffffu, sb, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float sa = up ? a4[i] : a4[i + 2], ka = up ? a4[i + 2] : a4[i];
            const float sb = up ? b4[i] : b4[i + 2], kb = up ? b4[i + 2] : b4[i];
            a2[i] = ka + __shfl_xor_sync(0xffffffffu, sa, 4);
            b2[i] = kb + __shfl_xor_sync(0xffffffffu, sb, 4);
        }
    }
    const bool up = lane & 2;
    const float sa = up ? a2[0] : a2[1], ka = up ? a2[1] : a2[0];
    const float sb = up ? b2[0] : b2[1], kb = up ? b2[1] : b2[0];
    ra = ka + __shfl_xor_sync(0xffffffffu, sa, 2);
    rb = kb + __shfl_xor_sync(0xffffffffu, sb, 2);
    ra += __shfl_xor_sync(0xffffffffu, ra, 1);
    rb += __shfl_xor_sync(0xffffffffu, rb, 1);
}

__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(TC_GROUP_THREADS) : "memory");
}

// MMA completion: only the group's first warp polls the mbarrier (try_wait parks the warp for a while per attempt),
// the other three sleep in the group's hardware barrier instead of spinning -- a quarter of the polling instructions
// (SYNCS + BRA + YIELD were 16 % of everything the kernel executed).
#ifndef WW_TC_WAIT1
#define WW_TC_WAIT1 1
#endif
#ifndef WW_TC_PREFETCH
#define WW_TC_PREFETCH 1
#endif
#ifndef WW_TC_CMVN2
#define WW_TC_CMVN2 1   // python CMVN of a warp's two windows in lock-step (tc_cmvn_py2)
#endif
__device__ __forceinline__ void tc_wait_mma(uint64_t* bar, uint32_t phase, int q4, int group) {
#if WW_TC_WAIT1
    if (q4 == 0) mbar_wait(bar, phase);
    group_sync(group);
#else
    mbar_wait(bar, phase);
#endif
}

// raw features of one window held by a warp: lane <-> frame (t = lane and t = lane + 32)
struct TcWin {
    float x0[WW_N_MFCC], x1[WW_N_MFCC];
};

// Where the CNN role runs.  Stand-alone: one CTA per SM over a feature batch in HBM.  FUSED (ww_fused.cuh): a few CTAs
// of the clip kernel, fed by the frontend pipelines of the same launch through an L2-resident ring of [13][63] windows
// (wait_ready / window / release), and the windows inside the guard band are copied out of the ring for the exact kernel.
struct TcSolo {
    static constexpr bool FUSED = false;
    __device__ __forceinline__ long long cta() const { return blockIdx.x; }
    __device__ __forceinline__ long long n_cta() const { return gridDim.x; }
    __device__ __forceinline__ void wait_ready(long long, long long, int) const {}
    __device__ __forceinline__ void release(long long) const {}
    __device__ __forceinline__ const float* window(long long) const { return nullptr; }
};

// ring windows: contiguous [13][63], written by other SMs during this launch -> L2 loads only (L1 is not coherent)
__device__ __forceinline__ void tc_load_window_ring(const float* wbase, bool live, int lane, TcWin& w) {
    const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
    const float* p0 = wbase + lane;
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) {
        w.x0[q] = live ? __ldcg(p0 + q * WW_WINDOW_FRAMES) : 0.f;
        w.x1[q] = (live && has1) ? __ldcg(p0 + q * WW_WINDOW_FRAMES + 32) : 0.f;
    }
}

// flat [n][13][63] batches: the 26 loads of a lane are `base + lane + constant`, i.e. immediate offsets of one address
// (the strided form below spends two IADD3 per load on its 64-bit pointers: 52 of S0's ~480 instructions per window),
// and a window past the end is skipped by a warp-uniform branch instead of 26 predicated selects
__device__ __forceinline__ void tc_load_window_flat(const float* wbase, bool live, int lane, TcWin& w) {
    const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
    if (live) {
        const float* p = wbase + lane;
#pragma unroll
        for (int q = 0; q < WW_N_MFCC; ++q) {
            w.x0[q] = p[q * WW_WINDOW_FRAMES];
            w.x1[q] = has1 ? p[q * WW_WINDOW_FRAMES + 32] : 0.f;
        }
    } else {
#pragma unroll
        for (int q = 0; q < WW_N_MFCC; ++q) w.x0[q] = w.x1[q] = 0.f;
    }
}

template <class ARGS>
__device__ __forceinline__ void tc_load_window(const ARGS& a, long long win, int lane, TcWin& w) {
    const bool live = win < a.n_windows;
    const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
    const float* wbase = a.feats;
    if (live)
        wbase = a.group_windows ? a.feats + (win / a.group_windows) * a.group_stride + (win % a.group_windows) * a.win_stride
                                : a.feats + win * a.win_stride;
    const float* p0 = wbase + lane * a.frame_stride;
    const float* p1 = p0 + 32 * a.frame_stride;
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) {
        w.x0[q] = live ? *p0 : 0.f;
        w.x1[q] = (live && has1) ? *p1 : 0.f;
        p0 += a.coef_stride;
        p1 += a.coef_stride;
    }
}

// device-style CMVN (esp_wake_word_detector.cpp:128-131,179-211): int8 rounding, population std, int8 output handed
// to the model at exponent -4.  EXACTLY the arithmetic of cnn_fp32_kernel so that all paths quantise identically;
// fully unrolled: 13 independent reduction chains interleave.  Leaves z = k/16 (k the int8 model input) in w.
__device__ __forceinline__ void tc_cmvn_device(TcWin& w, int lane) {
    const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) {
        const float v0 = lround_clamp_i8(w.x0[q]);
        const float v1 = has1 ? lround_clamp_i8(w.x1[q]) : 0.f;
        // v0, v1 are integers in [-128, 127]: their sum over the window is exact in any order, so the integer
        // warp reduction (one REDUX) gives the same float as cnn_fp32_kernel's shuffle tree
        const float mean = div63_exact((float)__reduce_add_sync(0xffffffffu, (int)v0 + (int)v1));
        const float d0 = v0 - mean, d1 = has1 ? v1 - mean : 0.f;
        const float ss = warp_sum(d0 * d0 + d1 * d1);   // float: keep the shuffle-tree order of cnn_fp32_kernel
        const float den = sqrtf(div63_exact(ss)) + 1e-8f;
        const float rinv = rcp_approx(den);
        w.x0[q] = fminf(fmaxf(div_lround_clamp_i8(d0, den, rinv) * 16.f, -128.f), 127.f) * 0.0625f;
        w.x1[q] = fminf(fmaxf(div_lround_clamp_i8(d1, den, rinv) * 16.f, -128.f), 127.f) * 0.0625f;
    }
}

// python-style CMVN of one window in place: two transposed warp reductions (16 shuffles each) instead of 26 butterflies
__device__ __forceinline__ void tc_cmvn_py(TcWin& w, int lane) {
    float(&x0)[WW_N_MFCC] = w.x0;
    float(&x1)[WW_N_MFCC] = w.x1;
    const float m1 = lane + 32 < WW_WINDOW_FRAMES ? 1.f : 0.f;
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = q < WW_N_MFCC ? x0[q] + x1[q] : 0.f;
    // lane 2q holds sum_t x[q][t]; the mean is a multiply here (the operand is rounded to fp16 anyway; the exact
    // division lives in cnn_fp32_kernel, which re-scores every window near the threshold)
    const float tot = reduce16(v, lane) * (1.f / (float)WW_WINDOW_FRAMES);
    float mean[WW_N_MFCC];
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) mean[q] = __shfl_sync(0xffffffffu, tot, 2 * q);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        if (q < WW_N_MFCC) {
            x0[q] -= mean[q];
            x1[q] = fmaf(-mean[q], m1, x1[q]);   // = has1 ? x1 - mean : 0 (x1 is 0 in lane 31): no select
            v[q] = fmaf(x0[q], x0[q], x1[q] * x1[q]);
        } else {
            v[q] = 0.f;
        }
    }
    const float ss = reduce16(v, lane);
    float sd = sqrtf(ss * (1.f / (float)(WW_WINDOW_FRAMES - 1)));
    if (sd == 0.f) sd = 1.f;
    const float inv = __frcp_rn(sd + 1e-8f);
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) {
        const float iq = __shfl_sync(0xffffffffu, inv, 2 * q);
        x0[q] *= iq;
        x1[q] *= iq;
    }
}

// The same for the warp's TWO windows in lock-step.  One window's CMVN is a chain of dependent shuffle levels (two
// reductions of five levels, two broadcasts: ~1000 cycles of latency with four warps per scheduler to hide it); written
// as two calls the compiler keeps the windows one after the other (each call branches on the CMVN mode).  Here every
// step is issued for both windows before the next one, so the two chains overlap.  Same operations in the same order
// per window: the bits do not change.
__device__ __forceinline__ void tc_cmvn_py2(TcWin& wa, TcWin& wb, int lane) {
    const float m1 = lane + 32 < WW_WINDOW_FRAMES ? 1.f : 0.f;
    float va[16], vb[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        va[q] = q < WW_N_MFCC ? wa.x0[q] + wa.x1[q] : 0.f;
        vb[q] = q < WW_N_MFCC ? wb.x0[q] + wb.x1[q] : 0.f;
    }
    float ta, tb;
    reduce16x2(va, vb, lane, ta, tb);
    ta *= (1.f / (float)WW_WINDOW_FRAMES);
    tb *= (1.f / (float)WW_WINDOW_FRAMES);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        if (q < WW_N_MFCC) {
            const float ma = __shfl_sync(0xffffffffu, ta, 2 * q), mb = __shfl_sync(0xffffffffu, tb, 2 * q);
            wa.x0[q] -= ma;
            wb.x0[q] -= mb;
            wa.x1[q] = fmaf(-ma, m1, wa.x1[q]);   // = has1 ? x1 - mean : 0 (x1 is 0 in lane 31): no select
            wb.x1[q] = fmaf(-mb, m1, wb.x1[q]);
            va[q] = fmaf(wa.x0[q], wa.x0[q], wa.x1[q] * wa.x1[q]);
            vb[q] = fmaf(wb.x0[q], wb.x0[q], wb.x1[q] * wb.x1[q]);
        } else {
            va[q] = 0.f;
            vb[q] = 0.f;
        }
    }
    float sa, sb;
    reduce16x2(va, vb, lane, sa, sb);
    float da = sqrtf(sa * (1.f / (float)(WW_WINDOW_FRAMES - 1))), db = sqrtf(sb * (1.f / (float)(WW_WINDOW_FRAMES - 1)));
    if (da == 0.f) da = 1.f;
    if (db == 0.f) db = 1.f;
    const float ia = __frcp_rn(da + 1e-8f), ib = __frcp_rn(db + 1e-8f);
#pragma unroll
    for (int q = 0; q < WW_N_MFCC; ++q) {
        const float qa = __shfl_sync(0xffffffffu, ia, 2 * q), qb = __shfl_sync(0xffffffffu, ib, 2 * q);
        wa.x0[q] *= qa;
        wa.x1[q] *= qa;
        wb.x0[q] *= qb;
        wb.x1[q] *= qb;
    }
}

// fp16 store of one normalised window into the conv1 operand rows 64*slot + t + 1
__device__ __forceinline__ void tc_store_a1(const TcWin& w, int slot, int lane, unsigned char* sA1) {
    const float(&x0)[WW_N_MFCC] = w.x0;
    const float(&x1)[WW_N_MFCC] = w.x1;
    // frame t -> parity tile t & 1, row 1 + 32*slot + (t >> 1); channels 0-7 -> chunk 0, 8-12 (+3 zeros) -> chunk 1
    {
        const float lo8[8] = {x0[0], x0[1], x0[2], x0[3], x0[4], x0[5], x0[6], x0[7]};
        const float hi8[8] = {x0[8], x0[9], x0[10], x0[11], x0[12], 0.f, 0.f, 0.f};
        unsigned char* dst = sA1 + (lane & 1) * A1_PAR + (1 + 32 * slot + (lane >> 1)) * 16;
        *reinterpret_cast<uint4*>(dst) = pack_h8(lo8);
        *reinterpret_cast<uint4*>(dst + A1_LBO) = pack_h8(hi8);
    }
    {
        // lane 31 holds the non-existent frame 63 (x1 == 0): it writes the zero pad row that ends the odd block
        const float lo8[8] = {x1[0], x1[1], x1[2], x1[3], x1[4], x1[5], x1[6], x1[7]};
        const float hi8[8] = {x1[8], x1[9], x1[10], x1[11], x1[12], 0.f, 0.f, 0.f};
        unsigned char* dst = sA1 + (lane & 1) * A1_PAR + (1 + 32 * slot + 16 + (lane >> 1)) * 16;
        *reinterpret_cast<uint4*>(dst) = pack_h8(lo8);
        *reinterpret_cast<uint4*>(dst + A1_LBO) = pack_h8(hi8);
    }
    if (slot == 0 && lane == 0) {
        // row 0 of the odd tile (x[-1] of the first window); the storage is shared with X3, so re-zero it every octet
        *reinterpret_cast<uint4*>(sA1 + A1_PAR) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sA1 + A1_PAR + A1_LBO) = make_uint4(0, 0, 0, 0);
    }
}

// CMVN (python or device style) of the warp's two windows and their fp16 store (slots `slot`, `slot + 1`)
__device__ __forceinline__ void tc_cmvn_store2(const TcArgs& a, TcWin& wa, TcWin& wb, int slot, int lane, unsigned char* sA1) {
    if (a.cmvn_mode == CMVN_PY) {
#if WW_TC_CMVN2
        tc_cmvn_py2(wa, wb, lane);
#else
        tc_cmvn_py(wa, lane);
        tc_cmvn_py(wb, lane);
#endif
    } else if (a.cmvn_mode == CMVN_DEVICE) {
        tc_cmvn_device(wa, lane);
        tc_cmvn_device(wb, lane);
    }
    tc_store_a1(wa, slot, lane, sA1);
    tc_store_a1(wb, slot + 1, lane, sA1);
}

template <class ROLE>
__device__ __forceinline__ void cnn_tc_body(const TcArgs& a, unsigned char* smem, const ROLE role) {
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TC_OFF_BAR + 32);
    float* sfc2 = reinterpret_cast<float*>(smem + TC_OFF_FC2);
    unsigned char* sW = smem + TC_OFF_W;

    // the warp index comes out of a shuffle so that the compiler knows it (and the group, the tile addresses and the TMEM
    // columns derived from it) to be warp-uniform: the UMMA descriptors are then built on the uniform datapath instead of
    // being moved there one MMA at a time (an ELECT / R2UR / BRA.U.ANY loop around each of the 44 MMAs of an octet)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = warp_index_uniform(tid);
    const int group = warp >> 2;     // independent 4-warp group (own octet stream, tiles, mbarrier, TMEM columns)
    const int q4 = warp & 3;         // TMEM lane quadrant this warp may read (= warp % 4)
    const int tig = tid & (TC_GROUP_THREADS - 1);
    const int C = a.num_classes;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR) + group;
    float* part = reinterpret_cast<float*>(smem + TC_OFF_PART) + group * (2 * 8 * 8);
    float* wnorm = reinterpret_cast<float*>(smem + TC_OFF_NORM) + group * (2 * 8);
    unsigned char* act = smem + TC_OFF_ACT + group * TC_ACT_BYTES;
    unsigned char* sA1 = act + TC_ACT_A1;
    unsigned char* sA2 = act + TC_ACT_A2;
    unsigned char* sX3 = act + TC_ACT_X3;
    unsigned char* sG = act + TC_ACT_G;

    // ---- one-time setup: zero the activation tiles, stage the weights, allocate TMEM ----
    for (int i = tid; i < (TC_SMEM - TC_OFF_ACT) / 16; i += TC_THREADS)
        reinterpret_cast<uint4*>(smem + TC_OFF_ACT)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < TC_W_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(a.wblob + i);
    for (int i = tid; i < C * 64; i += TC_THREADS) sfc2[i] = __ldg(a.fc2 + i);
    if (tid == 0) {
        for (int g = 0; g < TC_GROUPS; ++g) mbar_init(reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR) + g, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // programmatic dependent launch: everything above (tiles, weights, TMEM) ran while the frontend launch that produces
    // the features was finishing
    if constexpr (!ROLE::FUSED) pdl_wait();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0) + (uint32_t)(group * TC_GROUP_COLS);
    const uint32_t sA1a = smem_u32(sA1), sA2a = smem_u32(sA2), sX3a = smem_u32(sX3), sGa = smem_u32(sG);
    const uint32_t sWa = smem_u32(sW);
    uint32_t phase = 0;
    const uint32_t tlane = (uint32_t)(32 * q4) << 16;

    const long long n_oct = (a.n_windows + TC_CLIPS - 1) / TC_CLIPS;
    const long long oct_stride = role.n_cta() * TC_GROUPS;
    long long oct = role.cta() * TC_GROUPS + group;
    long long oct_held = -1;   // FUSED: octet whose ring slot this group still holds

    // each warp owns two windows of the octet (slots 2*q4, 2*q4 + 1); the other groups' GEMM stages and epilogues
    // hide the latency of these loads
    int np = 0;  // parity of this group's octet count: the norm slots are double-buffered (S0 of the next octet
                 // runs while warp 0 may still be in this octet's last epilogue)
    // [n][13][63] batches (one octet = 26 208 contiguous bytes, a multiple of 16 from a 16-byte aligned base): the
    // group's NEXT octet is pulled into L2 by one asynchronous bulk prefetch while this one is computed, so that the
    // feature loads of S0 wait for an L2 hit instead of HBM (stall_long_sb was 37 % of the samples, issue slots 39 %
    // busy: the kernel waits, it does not starve).  Sliding stream windows overlap 62/63 and are L2/L1-hot anyway.
    const bool flat = a.group_windows == 0 && a.frame_stride == 1 && a.coef_stride == WW_WINDOW_FRAMES &&
                      a.win_stride == WW_N_MFCC * WW_WINDOW_FRAMES && (reinterpret_cast<uintptr_t>(a.feats) & 15) == 0;
    auto prefetch_octet = [&](long long o) {
        if (!WW_TC_PREFETCH || ROLE::FUSED || !flat || o >= n_oct) return;
        long long wins = a.n_windows - o * TC_CLIPS;
        wins = wins < TC_CLIPS ? wins : TC_CLIPS;
        const uint32_t bytes = (uint32_t)(wins * WW_N_MFCC * WW_WINDOW_FRAMES * 4) & ~15u;
        if (bytes) bulk_prefetch_l2(a.feats + o * (long long)(TC_CLIPS * WW_N_MFCC * WW_WINDOW_FRAMES), bytes);
    };
#pragma unroll 1
    for (; oct < n_oct; oct += oct_stride, np ^= 1) {
        if (tig == 0) prefetch_octet(oct + oct_stride);   // one octet (~12 us of work) ahead; the first one is a cold load
        // ================= S0: CMVN two windows per warp, write A1 (fp16) =================
        {
            TcWin wa, wb;
            if constexpr (ROLE::FUSED) {
                role.wait_ready(oct, a.n_windows, lane);   // both blocks of all clips of the octet are in the ring
                const long long w0 = oct * TC_CLIPS + 2 * q4;
                tc_load_window_ring(role.window(w0), w0 < a.n_windows, lane, wa);
                tc_load_window_ring(role.window(w0 + 1), w0 + 1 < a.n_windows, lane, wb);
            } else if (flat) {
                const long long w0 = oct * TC_CLIPS + 2 * q4;
                const float* wbase = a.feats + w0 * (long long)(WW_N_MFCC * WW_WINDOW_FRAMES);
                tc_load_window_flat(wbase, w0 < a.n_windows, lane, wa);
                tc_load_window_flat(wbase + WW_N_MFCC * WW_WINDOW_FRAMES, w0 + 1 < a.n_windows, lane, wb);
            } else {
                tc_load_window(a, oct * TC_CLIPS + 2 * q4, lane, wa);
                tc_load_window(a, oct * TC_CLIPS + 2 * q4 + 1, lane, wb);
            }
            if (a.cmvn_mode == CMVN_NONE) {
                // the caller's features are fed as they are: their norm scales the guard band of these two windows
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int q = 0; q < WW_N_MFCC; ++q) {
                    sa = fmaf(wa.x0[q], wa.x0[q], fmaf(wa.x1[q], wa.x1[q], sa));
                    sb = fmaf(wb.x0[q], wb.x0[q], fmaf(wb.x1[q], wb.x1[q], sb));
                }
                sa = warp_sum(sa);
                sb = warp_sum(sb);
                if (lane == 0) {
                    wnorm[8 * np + 2 * q4] = sqrtf(sa);
                    wnorm[8 * np + 2 * q4 + 1] = sqrtf(sb);
                }
            }
            tc_cmvn_store2(a, wa, wb, 2 * q4, lane, sA1);
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);
        if constexpr (ROLE::FUSED) {
            // every warp of the group has passed the previous octet's last ring reads (S0 loads, re-score copy)
            if (tig == 0 && oct_held >= 0) role.release(oct_held);
            oct_held = oct;
        }

        // ================= conv1: 2 row tiles x {even, odd outputs} x 3 taps (K = 16) =================
        // even output 2j = W0.x[2j-1] + W1.x[2j] + W2.x[2j+1] -> taps (odd, R-1), (even, R), (odd, R)
        // odd  output 2j+1 = W0.x[2j] + W1.x[2j+1] + W2.x[2j+2] -> taps (even, R), (odd, R), (even, R+1)
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_f16(128, 32);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int par = 0; par < 2; ++par)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int src_par = par ? (r == 1) : (r != 1);             // tile holding this tap
                        const int shift = par ? (r == 2) : -(r == 0);              // row shift inside that tile
                        umma_f16(tmem + 64 * i + 32 * par,
                                 umma_desc_kmajor(sA1a + src_par * A1_PAR + (1 + 128 * i + shift) * 16, A1_LBO),
                                 umma_desc_kmajor(sWa + TC_W1 + r * W1_TAP, W1_LBO), idesc, r > 0);
                    }
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 1: lane = pooled position (w, j): max(even, odd, 0) over 32 channels -> conv2 operand ----
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
            float ve[32], vo[32];
            tmem_ld32x2(tmem + tlane + 64 * i, tmem + tlane + 64 * i + 32, ve, vo);
            const int g = 128 * i + 32 * q4 + lane;
            const int w = g >> 5, j = g & 31;
            float mine[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) mine[c] = fmaxf(fmaxf(ve[c], vo[c]), 0.f);
            if (j < 31) {
                unsigned char* dst = sA2 + (j & 1) * A2_PAR + (1 + 16 * w + (j >> 1)) * 16;
#pragma unroll
                for (int gch = 0; gch < 4; ++gch) *reinterpret_cast<uint4*>(dst + gch * A2_LBO) = pack_h8(mine + 8 * gch);
                if (a.dbg && oct == 0) {
                    // dbg[0 .. 8*31*32): pooled conv1 activations [clip][t][ch]
#pragma unroll
                    for (int c = 0; c < 32; ++c) a.dbg[(w * 31 + j) * 32 + c] = mine[c];
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= conv2: {even, odd outputs} x 3 taps x 2 K-steps (M = 128 pooled-pair rows) =================
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_f16(128, 64);
#pragma unroll
            for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int src_par = par ? (r == 1) : (r != 1);
                        const int shift = par ? (r == 2) : -(r == 0);
                        umma_f16(tmem + 64 * par,
                                 umma_desc_kmajor(sA2a + src_par * A2_PAR + (1 + shift) * 16 + ks * 2 * A2_LBO, A2_LBO),
                                 umma_desc_kmajor(sWa + TC_W2 + r * W2_TAP + ks * 2 * W2_LBO, W2_LBO), idesc,
                                 (r | ks) > 0);
                    }
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 2: lane = pooled position (w, m): 64 channels -> X3 row 1 + 16 w + m (natural order) ----
        {
            const int g = 32 * q4 + lane;
            const int w = g >> 4, m = g & 15;
            const bool valid = m < 15;
            unsigned char* dst = sX3 + (1 + g) * 16;
            if (g < 8) *reinterpret_cast<uint4*>(sX3 + g * X3_LBO) = make_uint4(0, 0, 0, 0);  // row 0 (shared with A1)
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                float ve[32], vo[32];
                tmem_ld32x2(tmem + tlane + 32 * hh, tmem + tlane + 64 + 32 * hh, ve, vo);
                float mine[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) mine[c] = valid ? fmaxf(fmaxf(ve[c], vo[c]), 0.f) : 0.f;
#pragma unroll
                for (int gch = 0; gch < 4; ++gch)
                    *reinterpret_cast<uint4*>(dst + (4 * hh + gch) * X3_LBO) = pack_h8(mine + 8 * gch);
                if (a.dbg && oct == 0 && valid) {
                    // dbg[8*31*32 ..): pooled conv2 activations [clip][t][ch]
                    float* d2 = a.dbg + 8 * 31 * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) d2[(w * 15 + m) * 64 + 32 * hh + c] = mine[c];
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= conv3 (channels on M): 3 taps x 4 K-steps, N = 128 positions =================
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_f16(128, 128);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_f16(tmem, umma_desc_kmajor(sWa + TC_W3 + r * W3_TAP + ks * 2 * W3_LBO, W3_LBO),
                             umma_desc_kmajor(sX3a + r * 16 + ks * 2 * X3_LBO, X3_LBO), idesc, (r | ks) > 0);
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 3: thread = channel o; per window ReLU + MaxPool + mean over the 7 pooled steps -> G ----
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int o = 32 * q4 + lane;
            float va[32], vb[32];
            tmem_ld32(tmem + tlane + 64 * h, va);
            tmem_ld32(tmem + tlane + 64 * h + 32, vb);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float* v = cc < 2 ? va + 16 * cc : vb + 16 * (cc - 2);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 7; ++j) s += fmaxf(fmaxf(v[2 * j], v[2 * j + 1]), 0.f);
                const float g = s * (1.f / 7.f);
                const int clip = 4 * h + cc;
                *reinterpret_cast<__half*>(sG + (o >> 3) * G_LBO + clip * 16 + (o & 7) * 2) = __float2half_rn(g);
                if (a.dbg && oct == 0) a.dbg[8 * 31 * 32 + 8 * 15 * 64 + clip * 128 + o] = g;
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(group);

        // ================= fc1: [128(64) x 128] . [16 x 128]^T, 8 K-steps =================
        if (tig == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_f16(128, 16);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                umma_f16(tmem, umma_desc_kmajor(sWa + TC_WF1 + ks * 2 * WF1_LBO, WF1_LBO),
                         umma_desc_kmajor(sGa + ks * 2 * G_LBO, G_LBO), idesc, ks > 0);
            umma_commit(bar);
        }
        tc_wait_mma(bar, phase, q4, group);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 4: ReLU, fc2 as a warp reduction (rows 0-63 = the warps of lane quadrants 0 and 1) ----
        if (q4 < 2) {
            float h[16];
            tmem_ld16(tmem + tlane, h);
            const int o = 32 * q4 + lane;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) h[c8] = fmaxf(h[c8], 0.f);
            if (a.dbg && oct == 0) {
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) a.dbg[8 * 31 * 32 + 8 * 15 * 64 + 8 * 128 + c8 * 64 + o] = h[c8];
            }
            for (int c = 0; c < C; ++c) {
                const float w = sfc2[c * 64 + o];
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const float s = warp_sum(h[c8] * w);
                    if (lane == 0) part[(q4 * 8 + c) * 8 + c8] = s;
                }
            }
        }
        tc_fence_before();
        group_sync(group);
        int resc_slot = -1;   // FUSED: position of this thread's window (tig = window-in-octet) in the re-score list
        if (tig < 8 * C) {
            const int c8 = tig & 7, c = tig >> 3;
            const long long win = oct * TC_CLIPS + c8;
            if (win < a.n_windows) {
                const float s = part[(0 * 8 + c) * 8 + c8] + part[(1 * 8 + c) * 8 + c8];
                a.logits[win * C + c] = s;
                if (c == 0 && a.decisions) {
                    unsigned char d = 0;
                    if (a.decide_mode == DECIDE_LOGIT) d = s > a.threshold;
                    else if (a.decide_mode == DECIDE_DEVICE) d = (1.f / (1.f + expf(-s)) * 100.f) >= a.threshold;
                    a.decisions[win] = d;
                }
                if (c == 0 && a.rescore_list) {
                    float bw = a.band;
                    if (a.cmvn_mode == CMVN_NONE) {
                        const float nx = wnorm[8 * np + c8];
                        bw = nx <= a.norm_limit ? fmaf(a.band_rel, nx, a.band) : __int_as_float(0x7f800000);
                    }
                    // negated comparisons: a NaN / Inf logit (fp16 overflow) is re-scored as well
                    bool near = !(fabsf(s - a.thr0) >= bw) || !(fabsf(s - a.thr1) >= bw);
                    if (C > 1) {  // the argmax over classes (CTC best path) must not depend on the operand precision
                        float top = s, second = -__int_as_float(0x7f800000);
                        for (int cc = 1; cc < C; ++cc) {
                            const float v = part[(0 * 8 + cc) * 8 + c8] + part[(1 * 8 + cc) * 8 + c8];
                            if (!(v <= top)) { second = top; top = v; }
                            else if (v > second) second = v;
                        }
                        near = near || !(top - second >= 2.f * bw);
                    }
                    if (near) {
                        const int slot = atomicAdd(a.rescore_count, 1);
                        a.rescore_list[slot] = win + (a.rescore_feat ? a.rescore_base : 0);
                        resc_slot = slot;
                    }
                }
            }
        }
        if (a.rescore_feat) {
            // the feature buffer is recycled before the re-score runs: the group's first warp copies the windows it just
            // listed (lanes 0..7 hold their list positions) into the compact buffer that the exact kernel reads
            if (q4 == 0) {
                unsigned m = __ballot_sync(0xffffffffu, resc_slot >= 0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int sl = __shfl_sync(0xffffffffu, resc_slot, src);
                    const long long w = oct * TC_CLIPS + src;
                    const float* from = ROLE::FUSED ? role.window(w) : a.feats + w * (long long)(WW_N_MFCC * WW_WINDOW_FRAMES);
                    float* to = a.rescore_feat + (long long)sl * (WW_N_MFCC * WW_WINDOW_FRAMES);
                    for (int i = lane; i < WW_N_MFCC * WW_WINDOW_FRAMES; i += 32) to[i] = __ldcg(from + i);
                }
            }
        }
        // `part` is rewritten only after the next octet's four group barriers
    }
    if constexpr (ROLE::FUSED) {
        group_sync(group);
        if (tig == 0 && oct_held >= 0) role.release(oct_held);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) cnn_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_launch_dependents();
    cnn_tc_body<TcSolo>(a, smem, TcSolo{});
}

// ---- stand-alone CMVN (normalize_mfcc / device CMVN over [n][13][63] windows) ------------------------------
// ww_cmvn's kernel: a warp per window, grid stride, lane <-> frame.  6 552 bytes per window (read + write): bound by
// HBM.  Python style is the exact formula of extract_mfcc.py:47-88 (true divisions, unbiased std, std == 0 -> 1);
// device style is tc_cmvn_device, i.e. bit-identical to what the CNN kernels feed their first layer.
struct CmvnArgs {
    const float* feats;   // [n][13][63]
    float* out;           // [n][13][63]
    long long n_windows;
    int cmvn_mode;
    // tc_load_window's view of the windows
    long long win_stride, coef_stride, frame_stride, group_windows, group_stride;
};

__global__ void __launch_bounds__(256) cmvn_rows_kernel(const CmvnArgs a) {
    const int lane = threadIdx.x & 31;
    const bool has1 = lane + 32 < WW_WINDOW_FRAMES;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long win = (long long)blockIdx.x * (blockDim.x >> 5) + warp_index_uniform(); win < a.n_windows; win += warps) {
        TcWin w;
        tc_load_window(a, win, lane, w);
        if (a.cmvn_mode == CMVN_PY) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = q < WW_N_MFCC ? w.x0[q] + w.x1[q] : 0.f;
            const float tot = reduce16(v, lane) / (float)WW_WINDOW_FRAMES;   // lane 2q: mean of coefficient q
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (q < WW_N_MFCC) {
                    const float mean = __shfl_sync(0xffffffffu, tot, 2 * q);
                    w.x0[q] -= mean;
                    w.x1[q] = has1 ? w.x1[q] - mean : 0.f;
                    v[q] = fmaf(w.x0[q], w.x0[q], w.x1[q] * w.x1[q]);
                } else {
                    v[q] = 0.f;
                }
            }
            const float ss = reduce16(v, lane);
            float sd = sqrtf(ss / (float)(WW_WINDOW_FRAMES - 1));
            if (sd == 0.f) sd = 1.f;
            const float den = sd + 1e-8f;
            const float rden = __frcp_rn(den);   // one correctly rounded reciprocal per row ...
#pragma unroll
            for (int q = 0; q < WW_N_MFCC; ++q) {
                // ... and one Markstein step per element: q0 = x r, q1 = q0 + fma(-q0, d, x) r is the correctly rounded
                // x / d (up to the rare double-rounding cases of the method: a last-bit matter inside the 2e-5 of this
                // call's parity bar; the CNN kernels do not use this path)
                const float dq = __shfl_sync(0xffffffffu, den, 2 * q), rq = __shfl_sync(0xffffffffu, rden, 2 * q);
                const float a0 = w.x0[q] * rq, a1 = w.x1[q] * rq;
                w.x0[q] = fmaf(fmaf(-a0, dq, w.x0[q]), rq, a0);
                w.x1[q] = fmaf(fmaf(-a1, dq, w.x1[q]), rq, a1);
            }
        } else if (a.cmvn_mode == CMVN_DEVICE) {
            tc_cmvn_device(w, lane);
        }
        float* o = a.out + win * (long long)(WW_N_MFCC * WW_WINDOW_FRAMES) + lane;
#pragma unroll
        for (int q = 0; q < WW_N_MFCC; ++q) {
            o[q * WW_WINDOW_FRAMES] = w.x0[q];
            if (has1) o[q * WW_WINDOW_FRAMES + 32] = w.x1[q];
        }
    }
}

// ---- host: fp16 weight blob in UMMA K-major layout ---------------------------------------------------
// element (row n, k) of a [rows x K] operand: byte (k/8)*rows*16 + n*16 + (k%8)*2
inline void tc_pack_operand(__half* dst, int rows, int K_padded, const float* src, int src_rows, int src_K,
                            long long row_stride, long long k_stride) {
    for (int kc = 0; kc < K_padded / 8; ++kc)
        for (int n = 0; n < rows; ++n)
            for (int i = 0; i < 8; ++i) {
                const int k = kc * 8 + i;
                const float v = (n < src_rows && k < src_K) ? src[n * row_stride + k * k_stride] : 0.f;
                dst[((size_t)kc * rows + n) * 8 + i] = __float2half_rn(v);
            }
}

inline void tc_build_blob(std::vector<unsigned char>& blob, const float* conv1, const float* conv2, const float* conv3,
                          const float* fc1) {
    blob.assign(TC_W_BYTES, 0);
    for (int r = 0; r < 3; ++r) {
        // torch conv weight [O][I][3]: element (o, i, r) at o*I*3 + i*3 + r
        tc_pack_operand(reinterpret_cast<__half*>(blob.data() + TC_W1 + r * W1_TAP), 32, 16, conv1 + r, 32, 13, 13 * 3, 3);
        tc_pack_operand(reinterpret_cast<__half*>(blob.data() + TC_W2 + r * W2_TAP), 64, 32, conv2 + r, 64, 32, 32 * 3, 3);
        tc_pack_operand(reinterpret_cast<__half*>(blob.data() + TC_W3 + r * W3_TAP), 128, 64, conv3 + r, 128, 64, 64 * 3, 3);
    }
    tc_pack_operand(reinterpret_cast<__half*>(blob.data() + TC_WF1), 64, 128, fc1, 64, 128, 128, 1);
}

}  // namespace ww
