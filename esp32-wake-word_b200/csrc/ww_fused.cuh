// One persistent kernel for the whole clip path (sm_100a):  PCM -> MFCC -> CMVN -> CNN (tcgen05) -> logits / decision.
//
// Replaces, per batch of 1 s clips, the chain  extract_features (ml_models/src/extract_mfcc.py:151-176) ->
// normalize_mfcc (:47-88) -> LightweightKWS.forward (wakeModel.py:29-34) -> sigmoid > 0.5 (main.py:52-53) that the
// chunked path (ww_api.cu:score_clips_dev) runs as frontend launch -> feature scratch in HBM -> CNN launch.
// Selected with ww_set_option(WW_OPT_FUSED, 2); same bits as the chunked path (tests/test_gpu_fused.py).
//
// The grid is one 512-thread CTA per SM, and the SMs are specialised:
//   * CTAs [0, n_cnn)        run the CNN role   = cnn_tc_body (ww_cnn_tc.cuh): four 4-warp groups, tcgen05 + TMEM
//   * CTAs [n_cnn, gridDim)  run TWO frontend pipelines each = mfcc_body (ww_mfcc.cuh), 256 threads and ~112 KB of shared
//     memory per pipeline, each with its own named barrier -- the same occupancy as two frontend CTAs per SM
// The features never go to HBM: a frontend pipeline writes the [13][32] half of a clip into slot (clip mod R) of a ring
// of R windows that stays resident in the 126 MB L2 (R = 4096 clips = 13 MB; the PCM, read exactly once, is loaded
// with an evict-first policy), and the CNN groups pull whole octets of clips out of it:
//   ready[o mod R/8]  += 1 per finished block (release, by one thread after the pipeline's barrier)    -> 16 per octet
//   freed[o mod R/8]  += 1 when the group that scored octet o has done its last read of the slot (release)
// a group waits for ready == 16 (gen + 1) with acquire loads before its first feature load; a pipeline looks at
// freed >= gen before the DCT of a block stores into a slot (the look is a relaxed load issued one pipeline stage
// earlier, so its latency is hidden; the acquire spin runs only when that look was too early).  Windows whose logit
// falls inside the guard band of the fp16 operands are copied out of the ring into a compact buffer and listed; the
// exact kernel (cnn_fp32_kernel) re-scores that list after the launch.
// All waits are bounded: a wait that expires sets *err and every later wait returns at once, so a protocol fault
// ends the launch with an error code instead of a hang.
//
// Measured on B200 (262 144 clips, tools/time_fused.py; ncu: profiles/r2_fused_*): DRAM traffic 31.8 KB per clip = 0.99 x
// the algorithmic bytes (chunked path: 38.8 KB, 1.21 x) and two launches per 131 072 clips instead of three -- at
// 26.9 M clips/s against 30.0 M for the chunked launches, which therefore stay the default.  Where the 10 % go, each
// measured by switching one thing off (profiles/experiments/README.md): the frontend pipelines alone on 138 SMs run at
// 30.9 M clips/s (more than 138/148 of the stand-alone kernel: their output stays in L2); the CNN role's code in the same
// kernel costs 2.4 % (register allocation), the release fence of the hand-over 1.6 %, the CNN role RUNNING on ten other
// SMs 5 % (not its polling: memory-system interference), and the back-pressure of the ring 4 % -- the CNN groups idle a
// third of the time, yet a pipeline finds its slot still held in 12 % of the blocks, for any ring size from 1024 to
// 16 384 clips (the static octet -> group map makes the slowest group set the pace).
#pragma once
#include "ww_cnn_tc.cuh"
#include "ww_mfcc.cuh"

namespace ww {

constexpr int FUSED_THREADS = 512;
#ifndef WW_FUSED_RING_SHIFT
#define WW_FUSED_RING_SHIFT 9
#endif
constexpr int FUSED_RING_SHIFT = WW_FUSED_RING_SHIFT;     // log2 of the ring size in octets
constexpr int FUSED_RING_OCTETS = 1 << FUSED_RING_SHIFT;
constexpr int FUSED_RING_CLIPS = 8 * FUSED_RING_OCTETS;   // 4096 clips = 13 MB
static_assert((1 << FUSED_RING_SHIFT) == FUSED_RING_OCTETS, "ring shift");
constexpr int FUSED_WIN_FLOATS = WW_N_MFCC * WW_WINDOW_FRAMES;
#ifndef WW_FUSED_NAP_P
#define WW_FUSED_NAP_P 400
#endif
#ifndef WW_FUSED_NAP_C
#define WW_FUSED_NAP_C 200
#endif
constexpr unsigned FUSED_NAP_PRODUCER = WW_FUSED_NAP_P, FUSED_NAP_CONSUMER = WW_FUSED_NAP_C;   // ns between polls

struct FusedRing {
    float* ring;        // [FUSED_RING_CLIPS][13][63]
    int* ready;         // [FUSED_RING_OCTETS] blocks stored, monotonic over the launch
    int* freed;         // [FUSED_RING_OCTETS] octets consumed, monotonic over the launch
    int* err;           // 0, or the code of the first wait that expired
#ifdef WW_FUSED_STATS
    unsigned long long* stats;   // experiment build: cycles {producers waiting, consumers waiting, consumer groups total, pipelines total}
#endif
};

struct FusedArgs {
    MfccArgs mf;
    TcArgs tc;
    FusedRing r;
    int n_cnn;          // CTAs that run the CNN role
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Spin until *p >= want (acquire); bounded, see the header.  One opaque asm block executed by every lane of the calling
// warp (the lanes read the same word: one broadcast transaction): no C++-level control flow, no call, no single-lane
// branch -- a noinline helper behind `if (lane == 0)` cost the frontend its convergence guarantees (every shuffle got a
// divergent-warp slow path) and its uniform datapath.  ~2^22 polls of >= 64 ns before the wait gives up and sets *err.
// `seen` is an earlier (relaxed) look at the word: when it already satisfies the wait nothing is loaded at all.
// nap_ns: sleep between polls -- a frontend warp that spins takes issue slots from the pipeline it shares the SM with, and
// every poll is a request to the one L2 slice that holds the counter.
__device__ __forceinline__ void fused_wait_ge(const int* p, int want, int* err, int code, int seen, unsigned nap_ns) {
    asm volatile(
        "{\n"
        ".reg .pred pd, pe, pt;\n"
        ".reg .s32 v, e;\n"
        ".reg .u32 n;\n"
        "setp.ge.s32 pd, %4, %1;\n"
        "@pd bra WW_FW_DONE_%=;\n"
        "mov.u32 n, 0;\n"
        "WW_FW_LOOP_%=:\n"
        "ld.acquire.gpu.global.s32 v, [%0];\n"
        "setp.ge.s32 pd, v, %1;\n"
        "@pd bra WW_FW_DONE_%=;\n"
        "nanosleep.u32 %5;\n"
        "ld.relaxed.gpu.global.s32 e, [%2];\n"
        "setp.ne.s32 pe, e, 0;\n"
        "@pe bra WW_FW_DONE_%=;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 pt, n, 2097152;\n"
        "@pt bra WW_FW_LOOP_%=;\n"
        "st.relaxed.gpu.global.s32 [%2], %3;\n"
        "WW_FW_DONE_%=:\n"
        "}\n" ::"l"(p),
        "r"(want), "l"(err), "r"(code), "r"(seen), "r"(nap_ns)
        : "memory");
}

// One frontend pipeline of the CTA.  sub_ (which of the two) is the same for every lane of a warp and is obtained through
// a warp reduction, i.e. in a uniform register, so that what derives from it -- block ids, the shared-memory base, the
// barrier id -- stays on the uniform datapath as it does in the stand-alone kernel, where it derives from blockIdx.
// (Two template copies of the frontend, one per pipeline, were measured too: 21.7 against 26 M clips/s -- the two
// pipelines of an SM then run disjoint 36 KB loops and thrash the instruction caches.)
struct MfccFused {
    static constexpr bool FUSED = true;
    int sub_;
    long long first_, stride_;
    FusedRing r;
    uint64_t pol_;
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x & (MFCC_THREADS - 1); }
    __device__ __forceinline__ long long first() const { return first_; }
    __device__ __forceinline__ long long stride() const { return stride_; }
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + sub_), "n"(MFCC_THREADS) : "memory");
    }
    __device__ __forceinline__ long long out_slot(long long sig) const { return sig & (FUSED_RING_CLIPS - 1); }
    __device__ __forceinline__ uint64_t pcm_policy() const { return pol_; }
    __device__ __forceinline__ int free_probe(long long sig) const {
        return ld_relaxed_gpu(r.freed + ((sig >> 3) & (FUSED_RING_OCTETS - 1)));
    }
    // whole warp, every lane holds the probe's result (normally seen >= gen: the slot was freed long ago)
    __device__ __forceinline__ void wait_free(long long sig, int seen) const {
        const long long o = sig >> 3;
        const int gen = (int)(o >> FUSED_RING_SHIFT);
#ifdef WW_FUSED_STATS
        const long long t0 = clock64();
#endif
        fused_wait_ge(r.freed + (o & (FUSED_RING_OCTETS - 1)), gen, r.err, 1, seen, FUSED_NAP_PRODUCER);
#ifdef WW_FUSED_STATS
        if ((threadIdx.x & 31) == 0) {
            const unsigned long long dt = (unsigned long long)(clock64() - t0);
            atomicAdd(r.stats + 0, dt);
            atomicAdd(r.stats + 8 + 512 + 2 * blockIdx.x + sub_, dt);   // per pipeline: its DCT warps' waiting
            if (seen < gen) {
                atomicAdd(r.stats + 4, 1ull);                 // slow-path entries
                atomicMax(r.stats + 5, dt);                   // longest wait
                atomicMax(r.stats + 6, (unsigned long long)o);  // last octet that had to wait
                if (dt > 20000) atomicAdd(r.stats + 7, 1ull);  // waits > 10 us
            }
        }
#endif
    }
    // one thread, after the pipeline's barrier that follows the block's stores
    __device__ __forceinline__ void publish(long long sig) const {
        red_release_gpu(r.ready + ((sig >> 3) & (FUSED_RING_OCTETS - 1)), 1);
    }
};

struct TcFused {
    static constexpr bool FUSED = true;
    int cta_, n_cta_;
    FusedRing r;
    __device__ __forceinline__ long long cta() const { return cta_; }
    __device__ __forceinline__ long long n_cta() const { return n_cta_; }
    // whole warp: every lane polls the same word (one broadcast transaction) and orders its own later loads behind it
    __device__ __forceinline__ void wait_ready(long long oct, long long n_windows, int) const {
        long long wins = n_windows - oct * 8;
        wins = wins < 8 ? wins : 8;
        const int want = (int)(oct >> FUSED_RING_SHIFT) * 16 + 2 * (int)wins;
#ifdef WW_FUSED_STATS
        const long long t0 = clock64();
#endif
        fused_wait_ge(r.ready + (oct & (FUSED_RING_OCTETS - 1)), want, r.err, 2, -1, FUSED_NAP_CONSUMER);
#ifdef WW_FUSED_STATS
        if ((threadIdx.x & 31) == 0) atomicAdd(r.stats + 1, (unsigned long long)(clock64() - t0));
#endif
    }
    __device__ __forceinline__ void release(long long oct) const {
        red_release_gpu(r.freed + (oct & (FUSED_RING_OCTETS - 1)), 1);
    }
    __device__ __forceinline__ const float* window(long long win) const {
        return r.ring + (win & (FUSED_RING_CLIPS - 1)) * FUSED_WIN_FLOATS;
    }
};

template <typename TIN>
struct FusedSmem {
    static constexpr int SUB = (MfccSmem<TIN, MEL_PY>::TOTAL + 127) / 128 * 128;   // one frontend pipeline
    static constexpr int TOTAL = 2 * SUB > TC_SMEM ? 2 * SUB : TC_SMEM;
    static_assert(TOTAL <= 232448, "shared memory budget of one SM");
};

template <typename TIN>
__global__ void __launch_bounds__(FUSED_THREADS, 1) fused_clip_kernel(const __grid_constant__ FusedArgs fa) {
    extern __shared__ __align__(128) unsigned char smem[];
    if ((int)blockIdx.x < fa.n_cnn) {
        TcFused role;
        role.cta_ = (int)blockIdx.x;
        role.n_cta_ = fa.n_cnn;
        role.r = fa.r;
#ifdef WW_FUSED_STATS
        const long long t0 = clock64();
#endif
        cnn_tc_body<TcFused>(fa.tc, smem, role);
#ifdef WW_FUSED_STATS
        if ((threadIdx.x & 31) == 0) atomicAdd(fa.r.stats + 2, (unsigned long long)(clock64() - t0));
#endif
    } else {
        MfccFused pipe;
        pipe.sub_ = (int)__reduce_max_sync(0xffffffffu, threadIdx.x / MFCC_THREADS);
        pipe.first_ = 2LL * ((int)blockIdx.x - fa.n_cnn) + pipe.sub_;
        pipe.stride_ = 2LL * ((int)gridDim.x - fa.n_cnn);
        pipe.r = fa.r;
        pipe.pol_ = l2_policy_evict_first();
#ifdef WW_FUSED_STATS
        const long long t0 = clock64();
#endif
        mfcc_body<TIN, MEL_PY, true, MfccFused>(fa.mf, smem + pipe.sub_ * FusedSmem<TIN>::SUB, pipe);
#ifdef WW_FUSED_STATS
        if ((threadIdx.x & 31) == 0) atomicAdd(fa.r.stats + 3, (unsigned long long)(clock64() - t0));
        if ((threadIdx.x & (MFCC_THREADS - 1)) == 0) fa.r.stats[8 + 2 * blockIdx.x + pipe.sub_] += (unsigned long long)(clock64() - t0);
#endif
    }
}

}  // namespace ww
