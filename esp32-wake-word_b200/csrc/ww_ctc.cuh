// CTC best-path decoding, keyword match and CTC loss forward/backward (sm_100a), warp per utterance.
//
// Replaces  CTCKeywordDetector.ctc_greedy_decode     ml_models/test.py:201-217   (MODE_KEEP_REPEATS)
//           THCHS30Trainer.decode_predictions        ml_models/ctc.py:453-471    (MODE_COLLAPSE)
//           `keyword in decoded_text`                ml_models/test.py:189-194
//           nn.CTCLoss(...)(log_probs, targets, input_lengths, target_lengths)
//                                                    ml_models/test.py:89,111-112; ml_models/ctc.py:369,396
// All of these are streaming reductions over [T, C] rows: HBM-bound, no tensor-core work.
#pragma once
#include <math_constants.h>
#include "ww_common.cuh"

namespace ww {

enum { DECODE_KEEP_REPEATS = 0, DECODE_COLLAPSE = 1 };

struct GreedyArgs {
    const float* lp;       // lp[t*t_stride + b*b_stride + c]
    long long t_stride, b_stride;
    int T, B, C;
    const int* lengths;    // optional [B] valid frames per utterance
    int mode;
    int* labels;           // [B][T]
    int* out_len;          // [B]
    const int* keyword;    // optional [K] label ids
    int K;
    unsigned char* hits;   // optional [B]
    int pre_argmax;        // 1: labels[b][t] already holds the per-frame argmax (ctc_argmax_rows_kernel); compact in place
    int vec_ok;            // argmax kernel: rows are 16-byte aligned and C % 4 == 0
};

constexpr int CTC_WARPS = 4;

__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_greedy_kernel(const GreedyArgs a) {
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    // grid stride over the utterances (64 CTAs per SM): measured neutral to slightly better than one CTA per four
    // utterances at the keyword shape (2.95 against 2.82 G utterances/s, 2^20 x T = 63 x C = 3)
    const long long warps_total = (long long)gridDim.x * CTC_WARPS;
    for (long long b = (long long)blockIdx.x * CTC_WARPS + warp; b < a.B; b += warps_total) {
    const int Tb = a.lengths ? min(max(a.lengths[b], 0), a.T) : a.T;
    const float* base = a.lp + b * a.b_stride;
    int* lab = a.labels + b * (long long)a.T;
    int pos = 0;
    int carry = 0;  // argmax of the previous frame (blank before the first frame)
    for (int tb = 0; tb < Tb; tb += 32) {
        const int t = tb + lane;
        int idx = 0;
        if (a.C == 1) {
            // binary posterior in logit form: class 1 (keyword) iff logit > 0, else blank
            if (t < Tb) idx = base[(long long)t * a.t_stride] > 0.f ? 1 : 0;
        } else if (a.C <= 32) {
            // lane per frame
            if (t < Tb) {
                const float* row = base + (long long)t * a.t_stride;
                float best = row[0];
                for (int c = 1; c < a.C; ++c) {
                    const float v = row[c];
                    if (v > best) { best = v; idx = c; }
                }
            }
        } else if (a.pre_argmax) {
            // wide vocabulary: the argmax of every frame was computed by the bandwidth-bound row kernel into the label
            // buffer itself; compaction below never writes past the frames already read (pos <= tb)
            if (t < Tb) idx = lab[t];
        } else {
            // warp per frame, lanes stride the classes (coalesced); first index wins ties
            const int nt = min(32, Tb - tb);
            for (int tt = 0; tt < nt; ++tt) {
                const float* row = base + (long long)(tb + tt) * a.t_stride;
                float best = -CUDART_INF_F;
                int bi = 0x7fffffff;
                for (int c = lane; c < a.C; c += 32) {
                    const float v = row[c];
                    if (v > best || bi == 0x7fffffff) { best = v; bi = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == tt) idx = bi;
            }
        }
        int prev = __shfl_up_sync(0xffffffffu, idx, 1);
        if (lane == 0) prev = carry;
        bool keep = (t < Tb) && (idx != 0);
        if (a.mode == DECODE_COLLAPSE) keep = keep && (idx != prev);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) lab[pos + __popc(m & ((1u << lane) - 1u))] = idx;
        pos += __popc(m);
        carry = __shfl_sync(0xffffffffu, idx, 31);
    }
    // pad the tail so the label matrix is deterministic
    for (int i = pos + lane; i < a.T; i += 32) lab[i] = 0;
    if (lane == 0) a.out_len[b] = pos;
    if (a.hits) {
        __syncwarp();
        bool hit = (a.K == 0);
        for (int s = lane; s + a.K <= pos && a.K > 0; s += 32) {
            bool eq = true;
            for (int k = 0; k < a.K; ++k) eq = eq && (lab[s + k] == a.keyword[k]);
            hit = hit || eq;
        }
        hit = __any_sync(0xffffffffu, hit);
        if (lane == 0) a.hits[b] = hit ? 1 : 0;
    }
    }
}

// Keyword shapes (T <= 64, C <= 4: the reference's test.py decoder sees T = 63 frames of 3 classes, the clip path feeds
// 63-window utterances of one binary logit).  ctc_greedy_kernel walks such an utterance as a chain of dependent global
// round trips (lengths -> frames 0..31 -> frames 32..63 -> labels written and re-read for the keyword match), one
// utterance per warp at a time: latency-bound at a third of the HBM peak.  Here a lane takes frames `lane` and
// `lane + 32` at once, the rows of the warp's NEXT utterance are in flight while this one is compacted, and the labels
// are compacted in shared memory: stored once, coalesced, and matched against the keyword without touching HBM again.
constexpr int CTC_SHORT_T = 64;
constexpr int CTC_SHORT_KW = 8;   // keyword labels held in registers; longer keywords are read through L1

template <int C>
struct GreedyRows {
    float v0[C], v1[C];
    int Tb;
};

template <int C>
__device__ __forceinline__ void greedy_short_load(const GreedyArgs& a, long long b, int lane, GreedyRows<C>& r) {
    r.Tb = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) r.v0[c] = r.v1[c] = 0.f;
    if (b >= a.B) return;
    r.Tb = a.lengths ? min(max(__ldg(a.lengths + b), 0), a.T) : a.T;
    const float* base = a.lp + b * a.b_stride;
    // the length is only needed to MASK: the loads are bounded by T so that they do not wait for it
    if (lane < a.T) {
        const float* row = base + (long long)lane * a.t_stride;
#pragma unroll
        for (int c = 0; c < C; ++c) r.v0[c] = __ldcs(row + c);
    }
    if (lane + 32 < a.T) {
        const float* row = base + (long long)(lane + 32) * a.t_stride;
#pragma unroll
        for (int c = 0; c < C; ++c) r.v1[c] = __ldcs(row + c);
    }
}

template <int C>
__device__ __forceinline__ int greedy_short_argmax(const float (&v)[C]) {
    if (C == 1) return v[0] > 0.f ? 1 : 0;   // binary posterior in logit form (as in ctc_greedy_kernel)
    float best = v[0];
    int idx = 0;
#pragma unroll
    for (int c = 1; c < C; ++c)
        if (v[c] > best) { best = v[c]; idx = c; }
    return idx;
}

template <int C>
__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_greedy_short_kernel(const GreedyArgs a) {
    __shared__ int s_lab[CTC_WARPS][CTC_SHORT_T + CTC_SHORT_KW];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    int* sl = s_lab[warp];
    int kw[CTC_SHORT_KW];
#pragma unroll
    for (int k = 0; k < CTC_SHORT_KW; ++k) kw[k] = (a.hits && k < a.K) ? __ldg(a.keyword + k) : 0;
    const long long warps_total = (long long)gridDim.x * CTC_WARPS;
    long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    GreedyRows<C> cur, nxt;
    greedy_short_load<C>(a, b, lane, cur);
    for (; b < a.B; b += warps_total) {
        greedy_short_load<C>(a, b + warps_total, lane, nxt);
        const int Tb = cur.Tb;
        const int i0 = lane < Tb ? greedy_short_argmax<C>(cur.v0) : 0;
        const int i1 = lane + 32 < Tb ? greedy_short_argmax<C>(cur.v1) : 0;
        bool k0 = lane < Tb && i0 != 0, k1 = lane + 32 < Tb && i1 != 0;
        if (a.mode == DECODE_COLLAPSE) {
            int p0 = __shfl_up_sync(0xffffffffu, i0, 1), p1 = __shfl_up_sync(0xffffffffu, i1, 1);
            const int c31 = __shfl_sync(0xffffffffu, i0, 31);
            if (lane == 0) { p0 = 0; p1 = c31; }
            k0 = k0 && i0 != p0;
            k1 = k1 && i1 != p1;
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
        const unsigned lt = (1u << lane) - 1u;
        const int n0 = __popc(m0), pos = n0 + __popc(m1);
        sl[lane] = 0;
        sl[lane + 32] = 0;
        if (lane < CTC_SHORT_KW) sl[CTC_SHORT_T + lane] = 0;
        __syncwarp();
        if (k0) sl[__popc(m0 & lt)] = i0;
        if (k1) sl[n0 + __popc(m1 & lt)] = i1;
        __syncwarp();
        int* lab = a.labels + b * (long long)a.T;
        if (lane < a.T) lab[lane] = sl[lane];
        if (lane + 32 < a.T) lab[lane + 32] = sl[lane + 32];
        if (lane == 0) a.out_len[b] = pos;
        if (a.hits) {
            bool hit = (a.K == 0);
            if (a.K > 0 && a.K <= CTC_SHORT_KW) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int s = lane + 32 * h;
                    bool eq = s + a.K <= pos;
#pragma unroll
                    for (int k = 0; k < CTC_SHORT_KW; ++k) eq = eq && (k >= a.K || sl[s + k] == kw[k]);
                    hit = hit || eq;
                }
            } else if (a.K > 0) {
                for (int s = lane; s + a.K <= pos; s += 32) {
                    bool eq = true;
                    for (int k = 0; k < a.K; ++k) eq = eq && (sl[s + k] == __ldg(a.keyword + k));
                    hit = hit || eq;
                }
            }
            hit = __any_sync(0xffffffffu, hit);
            if (lane == 0) a.hits[b] = hit ? 1 : 0;
        }
        __syncwarp();   // s_lab is rewritten by the next utterance
        cur = nxt;
    }
}

// Wide-vocabulary best path, step 1: argmax of every (b, t) row -- one warp per row, grid stride over all B*T rows,
// 16-byte streaming loads; first index wins ties (torch.argmax).  T*C*4 bytes per utterance, HBM-bound.  A warp per
// UTTERANCE (ctc_greedy_kernel alone) walks its T rows one after the other and reaches 3.5 % of HBM bandwidth at
// B = 256, T = 801, C = 4096.
// LPR = lanes per row: a power of two in [8, 32] that covers the row with one 16-byte (or 4-byte) element per lane when it
// is short, so that narrow vocabularies (C = 64: 16 float4) keep all 32 lanes busy with 2-4 rows per warp
template <int LPR>
__global__ void __launch_bounds__(256) ctc_argmax_rows_kernel(const GreedyArgs a) {
    const int lane = threadIdx.x & 31;
    constexpr int lpr = LPR;
    constexpr int rpw = 32 / lpr;                 // rows per warp iteration
    const int sub = lane / lpr, l = lane % lpr;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)a.B * a.T;
    if constexpr (LPR < 32) {
        // short rows (one 16-byte element per lane at most): a warp iteration is 2-4 rows = 256-512 bytes, and with one
        // iteration in flight per warp the kernel waited for HBM latency, not bandwidth (0.37 of the peak at C = 64).
        // Four iterations' loads are issued before the first reduction; row -> (b, t) by 32-bit division when it fits.
        if (a.vec_ok && rows < (1LL << 31)) {
            constexpr int U = 4;
            const int n4 = a.C >> 2;
            const unsigned T = (unsigned)a.T;
            for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp_index_uniform()) * (rpw * U); r0 < rows;
                 r0 += warps * (rpw * U)) {
                float4 v[U];
                unsigned bb[U], tt[U];
                bool live[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long r = r0 + u * rpw + sub;
                    const bool in = r < rows;
                    const unsigned ru = in ? (unsigned)r : 0u;
                    bb[u] = ru / T;
                    tt[u] = ru - bb[u] * T;
                    live[u] = in && l < n4 && !(a.lengths && (int)tt[u] >= min(max(a.lengths[bb[u]], 0), a.T));
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live[u])
                        v[u] = __ldcs(reinterpret_cast<const float4*>(a.lp + (long long)bb[u] * a.b_stride + (long long)tt[u] * a.t_stride) + l);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float best = -CUDART_INF_F;
                    int bi = 0x7fffffff;
                    if (live[u]) {
                        const int c = 4 * l;
                        best = v[u].x; bi = c;
                        if (v[u].y > best) { best = v[u].y; bi = c + 1; }
                        if (v[u].z > best) { best = v[u].z; bi = c + 2; }
                        if (v[u].w > best) { best = v[u].w; bi = c + 3; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        if (o < lpr) {
                            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                        }
                    }
                    const long long r = r0 + u * rpw + sub;
                    const bool rowlive = r < rows && !(a.lengths && (int)tt[u] >= min(max(a.lengths[bb[u]], 0), a.T));
                    if (rowlive && l == 0) a.labels[(long long)bb[u] * a.T + tt[u]] = bi;
                }
            }
            return;
        }
    }
    for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp_index_uniform()) * rpw; r0 < rows; r0 += warps * rpw) {
        const long long r = r0 + sub;
        const bool in = r < rows;
        const long long b = in ? r / a.T : 0;
        const int t = (int)(in ? r - b * a.T : 0);
        const bool live = in && !(a.lengths && t >= min(max(a.lengths[b], 0), a.T));
        const float* row = a.lp + b * a.b_stride + (long long)t * a.t_stride;
        float best = -CUDART_INF_F;
        int bi = 0x7fffffff;
        if (live) {
            if (a.vec_ok) {
                const float4* r4 = reinterpret_cast<const float4*>(row);
                const int n4 = a.C >> 2;
                for (int c4 = l; c4 < n4; c4 += lpr) {
                    const float4 v = __ldcs(r4 + c4);
                    const int c = 4 * c4;
                    if (v.x > best || bi == 0x7fffffff) { best = v.x; bi = c; }
                    if (v.y > best) { best = v.y; bi = c + 1; }
                    if (v.z > best) { best = v.z; bi = c + 2; }
                    if (v.w > best) { best = v.w; bi = c + 3; }
                }
            } else {
                for (int c = l; c < a.C; c += lpr) {
                    const float v = row[c];
                    if (v > best || bi == 0x7fffffff) { best = v; bi = c; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // segmented butterfly: partners stay inside the row's lane group
            if (o < lpr) {                   // warp-uniform
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
        }
        if (live && l == 0) a.labels[b * (long long)a.T + t] = bi;
    }
}

// ------------------------------------------------------------------------------------------------
// CTC loss
// ------------------------------------------------------------------------------------------------
struct CtcLossArgs {
    const float* lp;        // lp[t*t_stride + b*b_stride + c]   (log-softmax rows)
    long long t_stride, b_stride;
    int T, B, C, S;         // S = padded target length (row stride of `targets`)
    const int* targets;     // [B][S]
    const int* in_len;      // [B]
    const int* tgt_len;     // [B]
    int blank;
    int zero_infinity;
    float* nll;             // [B] per-sample negative log likelihood (before any reduction)
    float* alpha;           // workspace [B][T][2S+1]
    const float* grad_out;  // [B] upstream gradient of nll_b (backward only)
    float* grad;            // grad[t*gt_stride + b*gb_stride + c] (backward only)
    long long gt_stride, gb_stride;
    float* meta;            // workspace [B][4]: {nll, rows that carry a gradient, target has repeated labels, -} (split backward)
    float* ab;              // workspace [B][T][2S+1]: alpha + beta (split backward; alpha itself stays intact for a second backward)
    int skip_fill;          // backward: grad rows were pre-filled with exp(lp)*go by ctc_grad_fill_kernel
    int fill_vec;           // fill kernel: 16-byte vector path is legal (alignment and C % 4 checked on the host)
};

__device__ __forceinline__ float lse3(float a, float b, float c) {
    float m = fmaxf(a, fmaxf(b, c));
    if (m == -CUDART_INF_F) m = 0.f;
    return logf(expf(a - m) + expf(b - m) + expf(c - m)) + m;
}
__device__ __forceinline__ float lse2(float a, float b) {
    float m = fmaxf(a, b);
    if (m == -CUDART_INF_F) return -CUDART_INF_F;
    return logf(expf(a - m) + expf(b - m)) + m;
}

// MUFU-based variants for the prefetching kernels (ex2.approx / lg2.approx: ~3e-7 absolute per call on sums in [1, 3],
// far below the fp32 resolution of alpha/beta themselves, whose magnitude reaches 1e3..1e4 at T = 801)
__device__ __forceinline__ float lse3f(float a, float b, float c) {
    float m = fmaxf(a, fmaxf(b, c));
    if (m == -CUDART_INF_F) m = 0.f;
    return __logf(__expf(a - m) + __expf(b - m) + __expf(c - m)) + m;
}
__device__ __forceinline__ float lse2f(float a, float b) {
    float m = fmaxf(a, b);
    if (m == -CUDART_INF_F) return -CUDART_INF_F;
    return __logf(__expf(a - m) + __expf(b - m)) + m;
}

// dynamic smem per warp: 2*Lp floats (double-buffered alpha/beta row), Lp = 2S+1 rounded up to 32,
// plus (backward) S floats label accumulators and 2*S ints (first-occurrence map, duplicate list)
__host__ __device__ inline int ctc_lp(int S) { return ((2 * S + 1 + 31) / 32) * 32; }

__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_loss_fwd_kernel(const CtcLossArgs a) {
    extern __shared__ float ctc_sm[];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    if (b >= a.B) return;
    const int Lp = ctc_lp(a.S);
    float* buf = ctc_sm + warp * 2 * Lp;
    const int Tb = min(max(a.in_len[b], 0), a.T);
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    float* al = a.alpha + b * (long long)a.T * Lw;
    const float NEG = -CUDART_INF_F;

    if (Tb == 0) {
        if (lane == 0) {
            float v = Sb == 0 ? 0.f : CUDART_INF_F;
            if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
            a.nll[b] = v;
        }
        return;
    }
    // t = 0
    for (int s = lane; s < Lp; s += 32) {
        float v = NEG;
        if (s == 0) v = base[a.blank];
        else if (s == 1 && L > 1) v = base[tgt[0]];
        buf[s] = v;
        if (s < L) al[s] = v;
    }
    __syncwarp();
    int cur = 0;
    for (int t = 1; t < Tb; ++t) {
        const float* row = base + (long long)t * a.t_stride;
        const float* prev = buf + cur * Lp;
        float* next = buf + (cur ^ 1) * Lp;
        for (int s = lane; s < Lp; s += 32) {
            float v = NEG;
            if (s < L) {
                const int lab = (s & 1) ? tgt[s >> 1] : a.blank;
                const float a0 = prev[s];
                const float a1 = s >= 1 ? prev[s - 1] : NEG;
                const bool skip = (s & 1) && s >= 3 && tgt[s >> 1] != tgt[(s >> 1) - 1];
                const float a2 = skip ? prev[s - 2] : NEG;
                v = lse3(a0, a1, a2) + row[lab];
                al[(long long)t * Lw + s] = v;
            }
            next[s] = v;
        }
        cur ^= 1;
        __syncwarp();
    }
    if (lane == 0) {
        const float* fin = buf + cur * Lp;
        float ll = fin[L - 1];
        if (L > 1) ll = lse2(ll, fin[L - 2]);
        float v = -ll;
        if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
        a.nll[b] = v;
    }
}

__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_loss_bwd_kernel(const CtcLossArgs a) {
    extern __shared__ float ctc_sm[];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    if (b >= a.B) return;
    const int Lp = ctc_lp(a.S);
    const int per_warp = 2 * Lp + 3 * a.S;
    float* buf = ctc_sm + warp * per_warp;       // beta double buffer
    float* acc = buf + 2 * Lp;                   // [S] per-label log-sum of alpha*beta
    int* first = reinterpret_cast<int*>(acc + a.S);  // [S] first occurrence of the same label
    int* dups = first + a.S;                         // [S] positions whose label occurred before
    const int Tb = min(max(a.in_len[b], 0), a.T);
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    const float* al = a.alpha + b * (long long)a.T * Lw;
    float* gbase = a.grad + b * a.gb_stride;
    const float NEG = -CUDART_INF_F;

    // recompute nll from alpha (unclamped by zero_infinity)
    float nll;
    {
        float ll = NEG;
        if (Tb > 0) {
            ll = al[(long long)(Tb - 1) * Lw + L - 1];
            if (L > 1) ll = lse2(ll, al[(long long)(Tb - 1) * Lw + L - 2]);
        } else if (Sb == 0) {
            ll = 0.f;
        }
        nll = -ll;
    }
    const float go = a.grad_out ? a.grad_out[b] : 1.f;
    const bool dead = (a.zero_infinity && nll == CUDART_INF_F);

    // rows past the input length (and everything for an infinite loss under zero_infinity) are zero
    const int t_live = dead ? 0 : Tb;
    if (!a.skip_fill) {
        for (int t = t_live; t < a.T; ++t) {
            float* grow = gbase + (long long)t * a.gt_stride;
            for (int c = lane; c < a.C; c += 32) grow[c] = 0.f;
        }
    }
    if (t_live == 0) return;

    // first-occurrence map of the target labels
    int ndup = 0;
    for (int i0 = 0; i0 < Sb; i0 += 32) {
        const int i = i0 + lane;
        int f = i;
        if (i < Sb) {
            const int li = tgt[i];
            for (int j = 0; j < i; ++j)
                if (tgt[j] == li) { f = j; break; }
            first[i] = f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, i < Sb && f != i);
        if (i < Sb && f != i) dups[ndup + __popc(m & ((1u << lane) - 1u))] = i;
        ndup += __popc(m);
    }
    __syncwarp();
    bool blank_in_tgt = false;
    for (int i0 = 0; i0 < Sb; i0 += 32) {
        const int i = i0 + lane;
        blank_in_tgt |= __any_sync(0xffffffffu, i < Sb && tgt[i] == a.blank);
    }

    int cur = 0;
    for (int t = Tb - 1; t >= 0; --t) {
        const float* row = base + (long long)t * a.t_stride;
        float* grow = gbase + (long long)t * a.gt_stride;
        const float* nxt = buf + cur * Lp;
        float* now = buf + (cur ^ 1) * Lp;
        float blank_m = NEG;
        // beta_t and alpha_t + beta_t
        for (int s = lane; s < Lp; s += 32) {
            float v = NEG;
            if (s < L) {
                const int lab = (s & 1) ? tgt[s >> 1] : a.blank;
                if (t == Tb - 1) {
                    v = (s == L - 1 || s == L - 2) ? row[lab] : NEG;
                } else {
                    const float b0 = nxt[s];
                    const float b1 = s + 1 < L ? nxt[s + 1] : NEG;
                    const bool skip = (s & 1) && s + 2 < L && tgt[s >> 1] != tgt[(s >> 1) + 1];
                    const float b2 = skip ? nxt[s + 2] : NEG;
                    v = lse3(b0, b1, b2) + row[lab];
                }
                const float ab = al[(long long)t * Lw + s] + v;
                if (s & 1) acc[s >> 1] = ab;
                else blank_m = lse2(blank_m, ab);
            }
            now[s] = v;
        }
        cur ^= 1;
        // blank total: log-sum-exp across lanes
        {
            float m = blank_m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float e = (m == NEG) ? 0.f : expf(blank_m - m);
            e = warp_sum(e);
            blank_m = (m == NEG) ? NEG : logf(e) + m;
        }
        // default gradient for every class: exp(lp)
        if (!a.skip_fill)
            for (int c = lane; c < a.C; c += 32) grow[c] = expf(row[c]) * go;
        __syncwarp();
        // fold repeated labels into their first occurrence (rare, sequential)
        if (lane == 0)
            for (int d = 0; d < ndup; ++d) acc[first[dups[d]]] = lse2(acc[first[dups[d]]], acc[dups[d]]);
        __syncwarp();
        for (int i = lane; i < Sb; i += 32) {
            if (first[i] == i) {
                const int c = tgt[i];
                if (c != a.blank) {
                    const float lpv = row[c];
                    grow[c] = (expf(lpv) - expf(acc[i] + nll - lpv)) * go;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            // a target label equal to the blank index is folded into the blank class as PyTorch does
            float tot = blank_m;
            if (blank_in_tgt)
                for (int i = 0; i < Sb; ++i)
                    if (first[i] == i && tgt[i] == a.blank) tot = lse2(tot, acc[i]);
            const float lpv = row[a.blank];
            grow[a.blank] = (expf(lpv) - expf(tot + nll - lpv)) * go;
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// Prefetching variants for 2S+1 <= 32*K (K <= 4): same arithmetic as ctc_loss_fwd_kernel / ctc_loss_bwd_kernel,
// but every lane keeps its K extended states (label, skip flag) in registers and the gathers that do NOT depend on
// the recursion -- lp[t][label] (and alpha[t][s] in the backward pass) -- are issued CTC_PF time steps ahead with
// 4-byte cp.async copies into a shared-memory ring.  The warp-per-utterance recursion is a dependent chain over T;
// without the prefetch each step waits for a global gather (1.9 us / step forward, 6.2 us backward at T = 801,
// C = 4096).  A register ring does not work: a warp has six scoreboards, so waiting for any shared-memory or MUFU
// result also waits for the outstanding prefetch loads (measured: 50 % long-scoreboard stalls); cp.async groups are
// tracked separately.
// ------------------------------------------------------------------------------------------------
constexpr int CTC_PF = 8;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int K>
__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_loss_fwd_pf_kernel(const CtcLossArgs a) {
    extern __shared__ float ctc_sm[];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    if (b >= a.B) return;
    constexpr int Lp = 32 * K;
    constexpr int PER_WARP = 2 * (Lp + 2) + CTC_PF * Lp;
    float* buf = ctc_sm + warp * PER_WARP + 2;       // two leading pad cells: prev[s-1], prev[s-2] need no branch
    float* ring = ctc_sm + warp * PER_WARP + 2 * (Lp + 2);   // [CTC_PF][Lp] gathered lp[t][label of state s]
    const int Tb = min(max(a.in_len[b], 0), a.T);
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    float* al = a.alpha + b * (long long)a.T * Lw;
    const float NEG = -CUDART_INF_F;

    if (Tb == 0) {
        if (lane == 0) {
            float v = Sb == 0 ? 0.f : CUDART_INF_F;
            if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
            a.nll[b] = v;
        }
        return;
    }
    // per-state flags live in ONE integer register (bit k: state is live, bit 8+k: the skip transition exists):
    // as bools they become long-lived predicates, ptxas runs out of the 7 predicate registers and spills them to
    // local memory, and the reload sits on the recursion's critical path behind the outstanding cp.async traffic
    int lab[K];
    unsigned flags = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane + 32 * k;
        const bool lv = s < L;
        lab[k] = (lv && (s & 1)) ? tgt[s >> 1] : a.blank;
        if (lv) flags |= 1u << k;
        if (lv && (s & 1) && s >= 3 && tgt[s >> 1] != tgt[(s >> 1) - 1]) flags |= 256u << k;
    }
    float* b0 = buf;
    float* b1 = buf + Lp + 2;
    if (lane < 2) {
        b0[-1 - lane] = NEG;
        b1[-1 - lane] = NEG;
    }
    // t = 0
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane + 32 * k;
        float v = NEG;
        if (s == 0) v = base[a.blank];
        else if (s == 1 && L > 1) v = base[tgt[0]];
        b0[s] = v;
        if (((flags >> k) & 1u)) al[s] = v;
    }
    __syncwarp();
    // ring slot (t - 1) % CTC_PF holds the gathered log-probs of step t; one cp.async group per step
#pragma unroll
    for (int p = 0; p < CTC_PF; ++p) {
        const int t = 1 + p;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (t < Tb && ((flags >> k) & 1u)) cp_async4(ring + p * Lp + lane + 32 * k, base + (long long)t * a.t_stride + lab[k]);
        cp_async_commit();
    }
    int cur = 0;
#pragma unroll 1
    for (int t = 1; t < Tb; ++t) {
        const int p = (t - 1) & (CTC_PF - 1);
        const float* prev = cur ? b1 : b0;
        float* next = cur ? b0 : b1;
        cp_async_wait<CTC_PF - 1>();   // the group of step t has landed (each lane reads only what it copied)
        float lpv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) lpv[k] = ((flags >> k) & 1u) ? ring[p * Lp + lane + 32 * k] : 0.f;
        // refill this slot for step t + CTC_PF (independent of the recursion)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int tn = t + CTC_PF;
            if (tn < Tb && ((flags >> k) & 1u)) cp_async4(ring + p * Lp + lane + 32 * k, base + (long long)tn * a.t_stride + lab[k]);
        }
        cp_async_commit();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane + 32 * k;
            float v = NEG;
            if (((flags >> k) & 1u)) {
                const float a0 = prev[s], a1 = prev[s - 1];
                const float a2 = ((flags >> (8 + k)) & 1u) ? prev[s - 2] : NEG;
                v = lse3f(a0, a1, a2) + lpv[k];
                al[(long long)t * Lw + s] = v;
            }
            next[s] = v;
        }
        cur ^= 1;
        __syncwarp();
    }
    cp_async_wait<0>();
    if (lane == 0) {
        const float* fin = cur ? b1 : b0;
        float ll = fin[L - 1];
        if (L > 1) ll = lse2(ll, fin[L - 2]);
        float v = -ll;
        if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
        a.nll[b] = v;
    }
}

template <int K>
__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_loss_bwd_pf_kernel(const CtcLossArgs a) {
    extern __shared__ float ctc_sm[];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    if (b >= a.B) return;
    constexpr int Lp = 32 * K;
    const int per_warp = 3 * (Lp + 2) + 2 * CTC_PF * Lp + 3 * a.S;
    float* buf = ctc_sm + warp * per_warp;           // beta double buffer, two trailing pad cells each
    float* lps = buf + 2 * (Lp + 2);                 // [Lp] lp[t][label of state s] of the current step
    float* ring_lp = lps + (Lp + 2);                 // [CTC_PF][Lp] prefetched lp[t][label]
    float* ring_al = ring_lp + CTC_PF * Lp;          // [CTC_PF][Lp] prefetched alpha[t][s]
    float* acc = ring_al + CTC_PF * Lp;              // [S] per-label log-sum of alpha*beta
    int* first = reinterpret_cast<int*>(acc + a.S);  // [S] first occurrence of the same label
    int* dups = first + a.S;                         // [S] positions whose label occurred before
    const int Tb = min(max(a.in_len[b], 0), a.T);
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    const float* al = a.alpha + b * (long long)a.T * Lw;
    float* gbase = a.grad + b * a.gb_stride;
    const float NEG = -CUDART_INF_F;

    float nll;
    {
        float ll = NEG;
        if (Tb > 0) {
            ll = al[(long long)(Tb - 1) * Lw + L - 1];
            if (L > 1) ll = lse2(ll, al[(long long)(Tb - 1) * Lw + L - 2]);
        } else if (Sb == 0) {
            ll = 0.f;
        }
        nll = -ll;
    }
    const float go = a.grad_out ? a.grad_out[b] : 1.f;
    const bool dead = (a.zero_infinity && nll == CUDART_INF_F);
    const int t_live = dead ? 0 : Tb;
    if (!a.skip_fill) {
        for (int t = t_live; t < a.T; ++t) {
            float* grow = gbase + (long long)t * a.gt_stride;
            for (int c = lane; c < a.C; c += 32) grow[c] = 0.f;
        }
    }
    if (t_live == 0) return;

    int ndup = 0;
    for (int i0 = 0; i0 < Sb; i0 += 32) {
        const int i = i0 + lane;
        int f = i;
        if (i < Sb) {
            const int li = tgt[i];
            for (int j = 0; j < i; ++j)
                if (tgt[j] == li) { f = j; break; }
            first[i] = f;
        }
        const unsigned m = __ballot_sync(0xffffffffu, i < Sb && f != i);
        if (i < Sb && f != i) dups[ndup + __popc(m & ((1u << lane) - 1u))] = i;
        ndup += __popc(m);
    }
    bool blank_in_tgt = false;
    for (int i0 = 0; i0 < Sb; i0 += 32) {
        const int i = i0 + lane;
        blank_in_tgt |= __any_sync(0xffffffffu, i < Sb && tgt[i] == a.blank);
    }
    int lab[K];
    unsigned flags = 0;   // bit k: state live, bit 8+k: skip transition (one register, not predicates: see the forward kernel)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane + 32 * k;
        const bool lv = s < L;
        lab[k] = (lv && (s & 1)) ? tgt[s >> 1] : a.blank;
        if (lv) flags |= 1u << k;
        if (lv && (s & 1) && s + 2 < L && tgt[s >> 1] != tgt[(s >> 1) + 1]) flags |= 256u << k;
    }
    // labels this lane patches: i = lane + 32 j (first occurrences that are not the blank index), else -1
    int patch_c[2];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = lane + 32 * j;
        patch_c[j] = (i < Sb && first[i] == i && tgt[i] != a.blank) ? tgt[i] : -1;
    }
    float* b0 = buf;
    float* b1 = buf + Lp + 2;
    if (lane < 2) {
        b0[Lp + lane] = NEG;
        b1[Lp + lane] = NEG;
    }
    __syncwarp();

#pragma unroll
    for (int p = 0; p < CTC_PF; ++p) {
        const int t = Tb - 1 - p;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (t >= 0 && ((flags >> k) & 1u)) {
                cp_async4(ring_lp + p * Lp + lane + 32 * k, base + (long long)t * a.t_stride + lab[k]);
                cp_async4(ring_al + p * Lp + lane + 32 * k, al + (long long)t * Lw + lane + 32 * k);
            }
        cp_async_commit();
    }
    int cur = 0;
#pragma unroll 1
    for (int t = Tb - 1; t >= 0; --t) {
        {
            {
                const int p = (Tb - 1 - t) & (CTC_PF - 1);
                float* grow = gbase + (long long)t * a.gt_stride;
                const float* nxt = cur ? b1 : b0;
                float* now = cur ? b0 : b1;
                cp_async_wait<CTC_PF - 1>();
                float lpv[K], alv[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    lpv[k] = ((flags >> k) & 1u) ? ring_lp[p * Lp + lane + 32 * k] : 0.f;
                    alv[k] = ((flags >> k) & 1u) ? ring_al[p * Lp + lane + 32 * k] : 0.f;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int tn = t - CTC_PF;
                    if (tn >= 0 && ((flags >> k) & 1u)) {
                        cp_async4(ring_lp + p * Lp + lane + 32 * k, base + (long long)tn * a.t_stride + lab[k]);
                        cp_async4(ring_al + p * Lp + lane + 32 * k, al + (long long)tn * Lw + lane + 32 * k);
                    }
                }
                cp_async_commit();
                // posterior mass of the blank class, summed directly: exp(alpha + beta + nll - lp) <= 1 for every state,
                // so no log-sum-exp (max butterfly + log) is needed on the way to the gradient
                float blank_e = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int s = lane + 32 * k;
                    float v = NEG;
                    if (((flags >> k) & 1u)) {
                        if (t == Tb - 1) {
                            v = (s == L - 1 || s == L - 2) ? lpv[k] : NEG;
                        } else {
                            const float c0 = nxt[s];
                            const float c1 = s + 1 < L ? nxt[s + 1] : NEG;
                            const float c2 = ((flags >> (8 + k)) & 1u) ? nxt[s + 2] : NEG;
                            v = lse3f(c0, c1, c2) + lpv[k];
                        }
                        const float ab = alv[k] + v;
                        if (s & 1) acc[s >> 1] = ab;
                        else blank_e += __expf(ab + nll - lpv[k]);
                        lps[s] = lpv[k];
                    }
                    now[s] = v;
                }
                cur ^= 1;
                blank_e = warp_sum(blank_e);
                if (!a.skip_fill) {
                    const float* row = base + (long long)t * a.t_stride;
                    for (int c = lane; c < a.C; c += 32) grow[c] = expf(row[c]) * go;
                }
                __syncwarp();
                if (ndup) {
                    if (lane == 0)
                        for (int d = 0; d < ndup; ++d) acc[first[dups[d]]] = lse2(acc[first[dups[d]]], acc[dups[d]]);
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // Sb <= 2 * K * 16 - 1 <= 63: at most two labels per lane
                    const int i = lane + 32 * j;
                    if (patch_c[j] >= 0) {
                        const float l = lps[2 * i + 1];
                        grow[patch_c[j]] = (__expf(l) - __expf(acc[i] + nll - l)) * go;
                    }
                }
                if (lane == 0) {
                    // a target label equal to the blank index is folded into the blank class as PyTorch does (rare:
                    // the scan over the labels runs only for utterances that contain one)
                    const float l = lps[0];
                    float tot = blank_e;
                    if (blank_in_tgt)
                        for (int i = 0; i < Sb; ++i)
                            if (first[i] == i && tgt[i] == a.blank) tot += __expf(acc[i] + nll - l);
                    grow[a.blank] = (__expf(l) - tot) * go;
                }
                __syncwarp();
            }
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Wide-vocabulary backward in two passes (C >= 64, 2S+1 <= 128).
//   1. ctc_beta_pf_kernel: the beta recursion alone -- the same prefetching chain as the forward pass, mirrored -- which
//      writes alpha[t][s] + beta[t][s] into a second block of the workspace (alpha stays intact: a graph that is kept
//      may run the backward again) and leaves {nll, live rows, repeated-label flag} per utterance beside it.  A step is one log-sum-exp and one store; the gradient work that sat
//      on the recursion's critical path (blank reduction, label patches: 1.7 us per step at T = 801) is gone from it.
//   2. ctc_grad_rows_kernel: every (t, b) row independently, one warp per row, grid stride: exp(lp) * go streamed with
//      16-byte loads / stores (the former fill pass), then the <= S + 1 classes that occur in the target are patched
//      with exp(lp) - sum over their states of exp(alpha + beta + nll - lp).  Bandwidth-bound, 2 T C 4 bytes per
//      utterance, no dependence between rows.
// Measured at T = 801, B = 256, C = 4096, S = 32 (ncu launch list): forward 0.32 ms, beta 0.53 ms, rows 1.57 ms against
// fill 1.09-1.27 + recursion-with-patches 1.2 ms: 104 k against 90.5 k seq/s in the same harness (+15 %).  The rows pass
// runs at 4.5 TB/s; the stream-only fill followed by a patches-only pass is slower (1.27 + 0.41 ms).
// ------------------------------------------------------------------------------------------------
// ALONE = true: the recursion runs BESIDE the forward pass (ww_ctc_loss_fwd with WW_CTC_BETA_IN_FWD, on a side stream), so
// it may not read alpha: it stores beta itself, leaves only the repeated-label flag in `meta`, and the rows kernel adds
// alpha + beta and derives nll / the live rows from alpha's last row.
template <int K, bool ALONE = false>
__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_beta_pf_kernel(const CtcLossArgs a) {
    extern __shared__ float ctc_sm[];
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * CTC_WARPS + warp;
    if (b >= a.B) return;
    constexpr int Lp = 32 * K;
    constexpr int PER_WARP = 2 * (Lp + 2) + 2 * CTC_PF * Lp;
    float* buf = ctc_sm + warp * PER_WARP;           // beta double buffer, two trailing pad cells each
    float* ring_lp = buf + 2 * (Lp + 2);             // [CTC_PF][Lp] prefetched lp[t][label]
    float* ring_al = ring_lp + CTC_PF * Lp;          // [CTC_PF][Lp] prefetched alpha[t][s]
    const int Tb = min(max(a.in_len[b], 0), a.T);
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    const float* al = a.alpha + b * (long long)a.T * Lw;
    float* abw = a.ab + b * (long long)a.T * Lw;
    const float NEG = -CUDART_INF_F;

    float nll = 0.f;
    if constexpr (!ALONE) {
        float ll = NEG;
        if (Tb > 0) {
            ll = al[(long long)(Tb - 1) * Lw + L - 1];
            if (L > 1) ll = lse2(ll, al[(long long)(Tb - 1) * Lw + L - 2]);
        } else if (Sb == 0) {
            ll = 0.f;
        }
        nll = -ll;
    }
    const bool dead = !ALONE && (a.zero_infinity && nll == CUDART_INF_F);
    const int t_live = dead ? 0 : Tb;
    bool dup = false;
    for (int i0 = 0; i0 < Sb; i0 += 32) {
        const int i = i0 + lane;
        bool d = false;
        if (i < Sb) {
            const int li = tgt[i];
            d = li == a.blank;                     // a label equal to the blank index is folded into the blank class
            for (int j = 0; j < i && !d; ++j) d = tgt[j] == li;
        }
        dup |= __any_sync(0xffffffffu, d);
    }
    if (lane == 0) {
        float* m = a.meta + 4 * b;
        if constexpr (!ALONE) {
            m[0] = nll;
            m[1] = __int_as_float(t_live);
        }
        m[2] = __int_as_float(dup ? 1 : 0);
    }
    if (t_live == 0) return;

    int lab[K];
    unsigned flags = 0;   // bit k: state live, bit 8+k: skip transition (one register, not predicates: see the forward kernel)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane + 32 * k;
        const bool lv = s < L;
        lab[k] = (lv && (s & 1)) ? tgt[s >> 1] : a.blank;
        if (lv) flags |= 1u << k;
        if (lv && (s & 1) && s + 2 < L && tgt[s >> 1] != tgt[(s >> 1) + 1]) flags |= 256u << k;
    }
    float* b0 = buf;
    float* b1 = buf + Lp + 2;
    if (lane < 2) {
        b0[Lp + lane] = NEG;
        b1[Lp + lane] = NEG;
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < CTC_PF; ++p) {
        const int t = Tb - 1 - p;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (t >= 0 && ((flags >> k) & 1u)) {
                cp_async4(ring_lp + p * Lp + lane + 32 * k, base + (long long)t * a.t_stride + lab[k]);
                if constexpr (!ALONE) cp_async4(ring_al + p * Lp + lane + 32 * k, al + (long long)t * Lw + lane + 32 * k);
            }
        cp_async_commit();
    }
    int cur = 0;
#pragma unroll 1
    for (int t = Tb - 1; t >= 0; --t) {
        const int p = (Tb - 1 - t) & (CTC_PF - 1);
        const float* nxt = cur ? b1 : b0;
        float* now = cur ? b0 : b1;
        cp_async_wait<CTC_PF - 1>();
        float lpv[K], alv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            lpv[k] = ((flags >> k) & 1u) ? ring_lp[p * Lp + lane + 32 * k] : 0.f;
            alv[k] = (!ALONE && ((flags >> k) & 1u)) ? ring_al[p * Lp + lane + 32 * k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int tn = t - CTC_PF;   // rows >= 8 steps away from the one written below
            if (tn >= 0 && ((flags >> k) & 1u)) {
                cp_async4(ring_lp + p * Lp + lane + 32 * k, base + (long long)tn * a.t_stride + lab[k]);
                if constexpr (!ALONE) cp_async4(ring_al + p * Lp + lane + 32 * k, al + (long long)tn * Lw + lane + 32 * k);
            }
        }
        cp_async_commit();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane + 32 * k;
            float v = NEG;
            if (((flags >> k) & 1u)) {
                if (t == Tb - 1) {
                    v = (s == L - 1 || s == L - 2) ? lpv[k] : NEG;
                } else {
                    const float c0 = nxt[s];
                    const float c1 = s + 1 < L ? nxt[s + 1] : NEG;
                    const float c2 = ((flags >> (8 + k)) & 1u) ? nxt[s + 2] : NEG;
                    v = lse3f(c0, c1, c2) + lpv[k];
                }
                abw[(long long)t * Lw + s] = alv[k] + v;   // alpha + beta (ALONE: beta, alv is 0)
            }
            now[s] = v;
        }
        cur ^= 1;
        __syncwarp();
    }
    cp_async_wait<0>();
}

constexpr int CTC_ROWS_WARPS = 8;

// FILL = false: the rows were filled by ctc_grad_fill_kernel (a pure stream at 6.2 TB/s); this kernel only patches.
template <bool FILL, bool ALONE = false>
__global__ void __launch_bounds__(CTC_ROWS_WARPS * 32) ctc_grad_rows_kernel(const CtcLossArgs a) {
    __shared__ float se_all[CTC_ROWS_WARPS][64];   // repeated labels only: per-state posterior mass of the odd states
    const int lane = threadIdx.x & 31, warp = warp_index_uniform();
    float* se = se_all[warp];
    const long long warps = (long long)gridDim.x * CTC_ROWS_WARPS;
    const long long rows = (long long)a.T * a.B;
    const int Lw = 2 * a.S + 1;
    for (long long r = (long long)blockIdx.x * CTC_ROWS_WARPS + warp; r < rows; r += warps) {
        const int t = (int)(r / a.B);
        const long long b = r - (long long)t * a.B;
        const float4 m = *reinterpret_cast<const float4*>(a.meta + 4 * b);
        float nll = m.x;
        int t_live = __float_as_int(m.y);
        const bool dup = __float_as_int(m.z) != 0;
        if constexpr (ALONE) {
            // the beta recursion ran beside the forward pass and knew nothing of alpha: nll and the live rows from
            // alpha's last row, exactly as ctc_beta_pf_kernel derives them
            const int Tb0 = min(max(a.in_len[b], 0), a.T);
            const int Sb0 = min(max(a.tgt_len[b], 0), a.S);
            const int L0 = 2 * Sb0 + 1;
            float ll = -CUDART_INF_F;
            if (Tb0 > 0) {
                const float* al_last = a.alpha + (b * (long long)a.T + (Tb0 - 1)) * Lw;
                ll = al_last[L0 - 1];
                if (L0 > 1) ll = lse2(ll, al_last[L0 - 2]);
            } else if (Sb0 == 0) {
                ll = 0.f;
            }
            nll = -ll;
            t_live = (a.zero_infinity && nll == CUDART_INF_F) ? 0 : Tb0;
        }
        const float* row = a.lp + b * a.b_stride + (long long)t * a.t_stride;
        float* grow = a.grad + b * a.gb_stride + (long long)t * a.gt_stride;
        const bool live = t < t_live;
        const float go = live ? (a.grad_out ? a.grad_out[b] : 1.f) : 0.f;
        if (!live) {
            if (FILL) {
                if (a.fill_vec) {
                    float4* g4 = reinterpret_cast<float4*>(grow);
                    for (int c = lane; c < (a.C >> 2); c += 32) __stcs(g4 + c, make_float4(0.f, 0.f, 0.f, 0.f));
                } else {
                    for (int c = lane; c < a.C; c += 32) grow[c] = 0.f;
                }
            }
            continue;
        }
        // what the patches need is requested BEFORE the row is streamed, so that its latency hides behind the fill:
        // alpha + beta of this lane's states (2S+1 <= 128: at most four per lane), their labels and the labels' log-probs
        const int Sb = min(max(a.tgt_len[b], 0), a.S);
        const int L = 2 * Sb + 1;
        const int* tgt = a.targets + b * (long long)a.S;
        const float* ab = a.ab + (b * (long long)a.T + t) * Lw;
        float vab[4], lpl[4];
        int lab[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s = lane + 32 * k;
            vab[k] = 0.f;
            lab[k] = a.blank;
            if (s < L) {
                vab[k] = ab[s];
                if constexpr (ALONE) vab[k] = a.alpha[(b * (long long)a.T + t) * Lw + s] + vab[k];   // alpha + beta
                if (s & 1) lab[k] = tgt[s >> 1];
            }
            lpl[k] = s < L ? row[lab[k]] : 0.f;
        }
        // exp(lp) * go for the whole row.  Ordinary stores (not streaming ones): the lines are still in L2 when the
        // patches overwrite single elements of them
        if (FILL) {
            if (a.fill_vec) {
                const float4* r4 = reinterpret_cast<const float4*>(row);
                float4* g4 = reinterpret_cast<float4*>(grow);
                const int n4 = a.C >> 2;
                for (int c = lane; c < n4; c += 32) {
                    const float4 v = __ldcs(r4 + c);
                    g4[c] = make_float4(expf(v.x) * go, expf(v.y) * go, expf(v.z) * go, expf(v.w) * go);   // __stcs: -4 %
                }
            } else {
                for (int c = lane; c < a.C; c += 32) grow[c] = expf(row[c]) * go;
            }
            __syncwarp();   // the patches below overwrite elements other lanes have just stored
        }
        float blank_e = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s = lane + 32 * k;
            if (s < L) {
                const float e = __expf(vab[k] + nll - lpl[k]);
                if (s & 1) {
                    if (!dup) grow[lab[k]] = (__expf(lpl[k]) - e) * go;
                    else se[s >> 1] = e;
                } else {
                    blank_e += e;
                }
            }
        }
        const float lb = row[a.blank];
        if (dup) {
            // repeated labels (or a label equal to the blank index): the first occurrence sums the mass of all of them,
            // in index order, so the result does not depend on the lane schedule
            __syncwarp();
            for (int i = lane; i < Sb; i += 32) {
                const int c = tgt[i];
                bool first = true;
                for (int j = 0; j < i && first; ++j) first = tgt[j] != c;
                if (c == a.blank) {
                    blank_e += se[i];
                } else if (first) {
                    float sum = se[i];
                    for (int j = i + 1; j < Sb; ++j)
                        if (tgt[j] == c) sum += se[j];
                    const float l = row[c];
                    grow[c] = (__expf(l) - sum) * go;
                }
            }
            __syncwarp();   // se is rewritten by the warp's next row
        }
        blank_e = warp_sum(blank_e);
        if (lane == 0) grow[a.blank] = (__expf(lb) - blank_e) * go;
    }
}

// ------------------------------------------------------------------------------------------------
// Short-target specialisation (S <= 3, i.e. at most 7 extended states -- the reference's keyword shapes
// T ~ 63, C = 3, S = 1..2 of ml_models/test.py): 8 lanes per utterance, 4 utterances per warp, one state per
// lane, neighbours through width-8 shuffles, no shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int CTC_SMALL_G = 8;

__device__ __forceinline__ int group_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_small_fwd_kernel(const CtcLossArgs a) {
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const int s = lane & 7;
    long long b = ((long long)blockIdx.x * CTC_WARPS + warp) * 4 + (lane >> 3);
    const bool live = b < a.B;
    if (!live) b = a.B - 1;
    const int Tb = live ? min(max(a.in_len[b], 0), a.T) : 0;
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    float* al = a.alpha + b * (long long)a.T * Lw;
    const float NEG = -CUDART_INF_F;
    const bool vs = s < L;
    const int lab = (vs && (s & 1)) ? tgt[s >> 1] : a.blank;
    const bool skip = vs && (s & 1) && s >= 3 && tgt[s >> 1] != tgt[(s >> 1) - 1];
    const int Tw = group_max(Tb);

    float alpha = NEG;
    if (Tb > 0) {
        if (s == 0) alpha = base[a.blank];
        else if (s == 1 && L > 1) alpha = base[lab];
        if (vs) al[s] = alpha;
    }
    for (int t = 1; t < Tw; ++t) {
        float a1 = __shfl_up_sync(0xffffffffu, alpha, 1, CTC_SMALL_G);
        float a2 = __shfl_up_sync(0xffffffffu, alpha, 2, CTC_SMALL_G);
        if (s < 1) a1 = NEG;
        if (!skip) a2 = NEG;
        if (t < Tb && vs) {
            alpha = lse3(alpha, a1, a2) + base[(long long)t * a.t_stride + lab];
            al[(long long)t * Lw + s] = alpha;
        }
    }
    const float e1 = __shfl_sync(0xffffffffu, alpha, L - 1, CTC_SMALL_G);
    const float e2 = __shfl_sync(0xffffffffu, alpha, L > 1 ? L - 2 : 0, CTC_SMALL_G);
    if (live && s == 0) {
        float v;
        if (Tb == 0) v = Sb == 0 ? 0.f : CUDART_INF_F;
        else v = -(L > 1 ? lse2(e1, e2) : e1);
        if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
        a.nll[b] = v;
    }
}

__global__ void __launch_bounds__(CTC_WARPS * 32) ctc_small_bwd_kernel(const CtcLossArgs a) {
    const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
    const int s = lane & 7;
    long long b = ((long long)blockIdx.x * CTC_WARPS + warp) * 4 + (lane >> 3);
    const bool live = b < a.B;
    if (!live) b = a.B - 1;
    const int Tb = live ? min(max(a.in_len[b], 0), a.T) : 0;
    const int Sb = min(max(a.tgt_len[b], 0), a.S);
    const int L = 2 * Sb + 1;
    const int Lw = 2 * a.S + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    const float* base = a.lp + b * a.b_stride;
    const float* al = a.alpha + b * (long long)a.T * Lw;
    float* gbase = a.grad + b * a.gb_stride;
    const float NEG = -CUDART_INF_F;
    const bool vs = s < L;
    const int lab = (vs && (s & 1)) ? tgt[s >> 1] : a.blank;
    const bool skip = vs && (s & 1) && s + 2 < L && tgt[s >> 1] != tgt[(s >> 1) + 1];

    float nll;
    {
        float ll = NEG;
        if (Tb > 0) {
            ll = al[(long long)(Tb - 1) * Lw + L - 1];
            if (L > 1) ll = lse2(ll, al[(long long)(Tb - 1) * Lw + L - 2]);
        } else if (Sb == 0) {
            ll = 0.f;
        }
        nll = -ll;
    }
    const float go = a.grad_out ? a.grad_out[b] : 1.f;
    const bool dead = a.zero_infinity && nll == CUDART_INF_F;
    const int Tl = dead ? 0 : Tb;        // rows that carry a gradient
    const int Tw = group_max(Tl);

    // rows without gradient
    if (live)
        for (int t = Tl; t < a.T; ++t)
            for (int c = s; c < a.C; c += CTC_SMALL_G) gbase[(long long)t * a.gt_stride + c] = 0.f;

    float beta = NEG;
    for (int t = Tw - 1; t >= 0; --t) {
        float b1 = __shfl_down_sync(0xffffffffu, beta, 1, CTC_SMALL_G);
        float b2 = __shfl_down_sync(0xffffffffu, beta, 2, CTC_SMALL_G);
        if (s + 1 >= L) b1 = NEG;
        if (!skip) b2 = NEG;
        const bool on = t < Tl;
        const float* row = base + (long long)t * a.t_stride;
        float ab = NEG;
        if (on && vs) {
            if (t == Tb - 1) beta = (s == L - 1 || s == L - 2) ? row[lab] : NEG;
            else beta = lse3(beta, b1, b2) + row[lab];
            ab = al[(long long)t * Lw + s] + beta;
        }
        // per class: log-sum-exp of alpha*beta over the states carrying that class (<= 7 states)
        for (int c0 = 0; c0 < a.C; c0 += CTC_SMALL_G) {
            const int c = c0 + s;
            float m = NEG;
#pragma unroll
            for (int q = 0; q < CTC_SMALL_G; ++q) {
                const float v = __shfl_sync(0xffffffffu, ab, q, CTC_SMALL_G);
                const int l = __shfl_sync(0xffffffffu, lab, q, CTC_SMALL_G);
                if (l == c && q < L) m = fmaxf(m, v);
            }
            float sum = 0.f;
#pragma unroll
            for (int q = 0; q < CTC_SMALL_G; ++q) {
                const float v = __shfl_sync(0xffffffffu, ab, q, CTC_SMALL_G);
                const int l = __shfl_sync(0xffffffffu, lab, q, CTC_SMALL_G);
                if (l == c && q < L && m != NEG) sum += expf(v - m);
            }
            if (live && on && c < a.C) {
                const float lpv = row[c];
                const float tot = (m == NEG) ? NEG : logf(sum) + m;
                gbase[(long long)t * a.gt_stride + c] = (expf(lpv) - expf(tot + nll - lpv)) * go;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Keyword shapes proper (S <= 3, C <= 8: T ~ 63, C = 3, S = 1..2 of ml_models/test.py:99-119): ONE THREAD per utterance.
// All 2S+1 <= 7 states, the <= 3 labels and the skip flags live in registers, so a time step is straight-line code with
// no shuffles and no idle lanes (the 8-lanes-per-utterance kernels above spend 8 lanes on <= 7 states and 16 shuffles
// per class on the gradient: ~2.2 k warp instructions per utterance against ~0.4 k here).  Consecutive threads take
// consecutive utterances, so the log-prob rows [T][B][C] are read and the gradient written as contiguous 4C-byte pieces,
// and alpha is kept TIME-MAJOR in the workspace, al[(t (2S+1) + s) B + b], so that a warp's alpha traffic is coalesced
// (the workspace layout is private to the forward / backward pair; both pick this kernel from (S, C) alone).
// log-sum-exp through MUFU ex2 / lg2 as in the prefetching kernels.
// ------------------------------------------------------------------------------------------------
constexpr int CTC_TINY_THREADS = 128;
constexpr int CTC_TINY_MAX_C = 8;

struct CtcTinyUtt {
    int Tb, Sb, L;
    int tg0, tg1, tg2;     // labels (blank where the target is shorter)
    bool sk3, sk5;         // state 3 / 5 may be entered from two states back
};

__device__ __forceinline__ CtcTinyUtt ctc_tiny_utt(const CtcLossArgs& a, long long b) {
    CtcTinyUtt u;
    u.Tb = min(max(a.in_len[b], 0), a.T);
    u.Sb = min(max(a.tgt_len[b], 0), a.S);
    u.L = 2 * u.Sb + 1;
    const int* tgt = a.targets + b * (long long)a.S;
    u.tg0 = u.Sb > 0 ? tgt[0] : a.blank;
    u.tg1 = u.Sb > 1 ? tgt[1] : a.blank;
    u.tg2 = u.Sb > 2 ? tgt[2] : a.blank;
    u.sk3 = u.Sb > 1 && u.tg1 != u.tg0;
    u.sk5 = u.Sb > 2 && u.tg2 != u.tg1;
    return u;
}

__global__ void __launch_bounds__(CTC_TINY_THREADS) ctc_tiny_fwd_kernel(const CtcLossArgs a) {
    const long long b = (long long)blockIdx.x * CTC_TINY_THREADS + threadIdx.x;
    if (b >= a.B) return;
    const CtcTinyUtt u = ctc_tiny_utt(a, b);
    const int Lw = 2 * a.S + 1;
    const float NEG = -CUDART_INF_F;
    const float* base = a.lp + b * a.b_stride;
    float* al = a.alpha + b;
    const long long B = a.B;
    float a0 = NEG, a1 = NEG, a2 = NEG, a3 = NEG, a4 = NEG, a5 = NEG, a6 = NEG;
    auto store = [&](int t) {
        float* row = al + (long long)t * Lw * B;
        row[0] = a0;
        if (Lw > 1) { row[B] = a1; row[2 * B] = a2; }
        if (Lw > 3) { row[3 * B] = a3; row[4 * B] = a4; }
        if (Lw > 5) { row[5 * B] = a5; row[6 * B] = a6; }
    };
    if (u.Tb > 0) {
        a0 = base[a.blank];
        if (u.L > 1) a1 = base[u.tg0];
        store(0);
    }
    for (int t = 1; t < u.Tb; ++t) {
        const float* row = base + (long long)t * a.t_stride;
        const float lb = row[a.blank];
        const float n0 = a0 + lb;
        float n1 = NEG, n2 = NEG, n3 = NEG, n4 = NEG, n5 = NEG, n6 = NEG;
        if (u.L > 1) {
            n1 = lse2f(a1, a0) + row[u.tg0];
            n2 = lse2f(a2, a1) + lb;
        }
        if (u.L > 3) {
            n3 = lse3f(a3, a2, u.sk3 ? a1 : NEG) + row[u.tg1];
            n4 = lse2f(a4, a3) + lb;
        }
        if (u.L > 5) {
            n5 = lse3f(a5, a4, u.sk5 ? a3 : NEG) + row[u.tg2];
            n6 = lse2f(a6, a5) + lb;
        }
        a0 = n0; a1 = n1; a2 = n2; a3 = n3; a4 = n4; a5 = n5; a6 = n6;
        store(t);
    }
    float v;
    if (u.Tb == 0) v = u.Sb == 0 ? 0.f : CUDART_INF_F;
    else if (u.L == 1) v = -a0;
    else if (u.L == 3) v = -lse2f(a2, a1);
    else if (u.L == 5) v = -lse2f(a4, a3);
    else v = -lse2f(a6, a5);
    if (a.zero_infinity && v == CUDART_INF_F) v = 0.f;
    a.nll[b] = v;
}

__global__ void __launch_bounds__(CTC_TINY_THREADS) ctc_tiny_bwd_kernel(const CtcLossArgs a) {
    const long long b = (long long)blockIdx.x * CTC_TINY_THREADS + threadIdx.x;
    if (b >= a.B) return;
    const CtcTinyUtt u = ctc_tiny_utt(a, b);
    const int Lw = 2 * a.S + 1, C = a.C;
    const float NEG = -CUDART_INF_F;
    const float* base = a.lp + b * a.b_stride;
    const float* al = a.alpha + b;
    float* gbase = a.grad + b * a.gb_stride;
    const long long B = a.B;

    float nll;
    {
        float ll = NEG;
        if (u.Tb > 0) {
            const float* row = al + (long long)(u.Tb - 1) * Lw * B;
            ll = row[(long long)(u.L - 1) * B];
            if (u.L > 1) ll = lse2f(ll, row[(long long)(u.L - 2) * B]);
        } else if (u.Sb == 0) {
            ll = 0.f;
        }
        nll = -ll;
    }
    const float go = a.grad_out ? a.grad_out[b] : 1.f;
    const bool dead = a.zero_infinity && nll == CUDART_INF_F;
    const int Tl = dead ? 0 : u.Tb;   // rows that carry a gradient
    for (int t = Tl; t < a.T; ++t) {
        float* g = gbase + (long long)t * a.gt_stride;
#pragma unroll
        for (int c = 0; c < CTC_TINY_MAX_C; ++c)
            if (c < C) g[c] = 0.f;
    }
    float b0 = NEG, b1 = NEG, b2 = NEG, b3 = NEG, b4 = NEG, b5 = NEG, b6 = NEG;
    for (int t = Tl - 1; t >= 0; --t) {
        const float* row = base + (long long)t * a.t_stride;
        float lpc[CTC_TINY_MAX_C];
#pragma unroll
        for (int c = 0; c < CTC_TINY_MAX_C; ++c) lpc[c] = c < C ? row[c] : 0.f;
        const float lb = row[a.blank], l0 = row[u.tg0], l1 = row[u.tg1], l2 = row[u.tg2];
        if (t == u.Tb - 1) {
            // only the last two states may end the path
            b0 = u.L == 1 ? lb : NEG;
            b1 = u.L == 3 ? l0 : NEG;
            b2 = u.L == 3 ? lb : NEG;
            b3 = u.L == 5 ? l1 : NEG;
            b4 = u.L == 5 ? lb : NEG;
            b5 = u.L == 7 ? l2 : NEG;
            b6 = u.L == 7 ? lb : NEG;
        } else {
            // beta[s] = lse(beta[s], beta[s+1], skip ? beta[s+2]) + lp[label(s)], states >= L stay -inf
            const float m0 = (u.L > 1 ? lse2f(b0, b1) : b0) + lb;
            float m1 = NEG, m2 = NEG, m3 = NEG, m4 = NEG, m5 = NEG, m6 = NEG;
            if (u.L > 1) {
                m1 = lse3f(b1, b2, u.sk3 ? b3 : NEG) + l0;
                m2 = (u.L > 3 ? lse2f(b2, b3) : b2) + lb;
            }
            if (u.L > 3) {
                m3 = lse3f(b3, b4, u.sk5 ? b5 : NEG) + l1;
                m4 = (u.L > 5 ? lse2f(b4, b5) : b4) + lb;
            }
            if (u.L > 5) {
                m5 = lse2f(b5, b6) + l2;
                m6 = b6 + lb;
            }
            b0 = m0; b1 = m1; b2 = m2; b3 = m3; b4 = m4; b5 = m5; b6 = m6;
        }
        // posterior mass of every state relative to its own class: exp(alpha + beta + nll - lp[label]) <= 1
        const float* arow = al + (long long)t * Lw * B;
        float eb = __expf(arow[0] + b0 + nll - lb), e1 = 0.f, e3 = 0.f, e5 = 0.f;
        if (u.L > 1) {
            e1 = __expf(arow[B] + b1 + nll - l0);
            eb += __expf(arow[2 * B] + b2 + nll - lb);
        }
        if (u.L > 3) {
            e3 = __expf(arow[3 * B] + b3 + nll - l1);
            eb += __expf(arow[4 * B] + b4 + nll - lb);
        }
        if (u.L > 5) {
            e5 = __expf(arow[5 * B] + b5 + nll - l2);
            eb += __expf(arow[6 * B] + b6 + nll - lb);
        }
        float* g = gbase + (long long)t * a.gt_stride;
#pragma unroll
        for (int c = 0; c < CTC_TINY_MAX_C; ++c) {
            if (c < C) {
                float res = c == a.blank ? eb : 0.f;
                if (u.Sb > 0 && c == u.tg0) res += e1;
                if (u.Sb > 1 && c == u.tg1) res += e3;
                if (u.Sb > 2 && c == u.tg2) res += e5;
                g[c] = (__expf(lpc[c]) - res) * go;
            }
        }
    }
}

// Wide-vocabulary backward, step 1: every (t, b) row gets exp(lp) * grad_out (or 0 past the input length / for
// an infinite loss under zero_infinity) from a fully parallel, bandwidth-bound pass -- one warp per row, grid
// stride.  ctc_loss_bwd_kernel (skip_fill = 1) then only patches the <= S+1 target classes of each row.
__global__ void __launch_bounds__(256) ctc_grad_fill_kernel(const CtcLossArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)a.T * a.B;
    const int Lw = 2 * a.S + 1;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp_index_uniform(); r < rows; r += warps) {
        const int t = (int)(r / a.B);
        const long long b = r - (long long)t * a.B;
        const int Tb = min(max(a.in_len[b], 0), a.T);
        float scale = 0.f;
        if (t < Tb) {
            scale = a.grad_out ? a.grad_out[b] : 1.f;
            if (a.zero_infinity) {
                const int Sb = min(max(a.tgt_len[b], 0), a.S);
                const int L = 2 * Sb + 1;
                const float* al = a.alpha + (b * (long long)a.T + (Tb - 1)) * Lw;
                float ll = al[L - 1];
                if (L > 1) ll = lse2(ll, al[L - 2]);
                if (ll == -CUDART_INF_F) scale = 0.f;
            }
        }
        const float* row = a.lp + b * a.b_stride + (long long)t * a.t_stride;
        float* grow = a.grad + b * a.gb_stride + (long long)t * a.gt_stride;
        if (a.fill_vec) {
            // rows are 16-byte aligned and C % 4 == 0: 16-byte streaming loads / stores
            const float4* r4 = reinterpret_cast<const float4*>(row);
            float4* g4 = reinterpret_cast<float4*>(grow);
            const int n4 = a.C >> 2;
            if (scale == 0.f) {
                for (int c = lane; c < n4; c += 32) __stcs(g4 + c, make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
                for (int c = lane; c < n4; c += 32) {
                    const float4 v = __ldcs(r4 + c);
                    __stcs(g4 + c, make_float4(expf(v.x) * scale, expf(v.y) * scale, expf(v.z) * scale, expf(v.w) * scale));
                }
            }
        } else if (scale == 0.f) {
            for (int c = lane; c < a.C; c += 32) grow[c] = 0.f;
        } else {
            for (int c = lane; c < a.C; c += 32) grow[c] = expf(row[c]) * scale;
        }
    }
}

}  // namespace ww
