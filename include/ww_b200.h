/*
 * ww_b200.h -- C ABI of libwwb200.so, the B200 (sm_100a) batch engine for the wake-word hot path
 *              MFCC -> CMVN -> LightweightKWS CNN -> decision / CTC scoring.
 *
 * Every entry point names the reference interface it stands in for (paths relative to the
 * Socrates666/esp32-wake-word tree).  Conventions:
 *   - opaque context, no hidden statics (the reference's mfcc.c:37-38,362-363 caches are not thread-safe)
 *   - device-buffer calls take caller-owned CUDA device pointers and an explicit cudaStream_t (as void*);
 *     they enqueue work and return without synchronising
 *   - host-buffer calls (`*_host`) take plain host pointers, do their own H2D/D2H and return when done
 *   - int status: WW_OK (0) or a negative WW_ERR_* (mirrors ESP_OK / ESP_FAIL,
 *     main/esp_wake_word_detector/include/esp_wake_word_detector.hpp:35); no exceptions cross the ABI
 *   - there is no CPU fallback: without a CUDA device ww_create() fails
 */
#ifndef WW_B200_H_
#define WW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ww_ctx ww_ctx;
typedef void* ww_stream_t; /* cudaStream_t */

enum {
    WW_OK = 0,
    WW_ERR_INVALID = -1,     /* bad argument (the reference returns NULL / ESP_FAIL, mfcc.c:434-437) */
    WW_ERR_CUDA = -2,        /* CUDA runtime error, see ww_last_error() */
    WW_ERR_NO_WEIGHTS = -3,  /* ww_load_weights() has not been called */
    WW_ERR_UNSUPPORTED = -4,
    WW_ERR_NOMEM = -5,
    WW_ERR_BUSY = -6         /* another thread is inside a fused call on this context (one at a time per context) */
};

/* feature definition */
enum {
    WW_FEAT_PY = 0,  /* torchaudio T.MFCC as configured at ml_models/src/extract_mfcc.py:137-148,171 */
    WW_FEAT_ESP = 1  /* main/esp_mfcc/mfcc.c:431-527 (extract_mfcc), intended math */
};
enum { WW_PCM_S16 = 0, WW_PCM_F32 = 1 };
/* feature output layout */
enum { WW_LAYOUT_COEF_MAJOR = 0 /* [n][13][T], PY */, WW_LAYOUT_FRAME_MAJOR = 1 /* [n][T][13], mfcc.c */ };
enum {
    WW_CMVN_NONE = 0,
    WW_CMVN_PY = 1,     /* normalize_mfcc(...,'cmvn'|'standardization'), extract_mfcc.py:47-88 */
    WW_CMVN_DEVICE = 2  /* detect_task, esp_wake_word_detector.cpp:128-131,179-211 */
};
enum {
    WW_DECIDE_NONE = 0,
    WW_DECIDE_LOGIT = 1,  /* logit > threshold; threshold 0 == sigmoid(out) > 0.5, ml_models/main.py:53 */
    WW_DECIDE_DEVICE = 2  /* 1/(1+expf(-x))*100 >= threshold (80), esp_wake_word_detector.cpp:226-228,245 */
};
enum {
    WW_DECODE_KEEP_REPEATS = 0, /* CTCKeywordDetector.ctc_greedy_decode, ml_models/test.py:201-217 */
    WW_DECODE_COLLAPSE = 1      /* THCHS30Trainer.decode_predictions, ml_models/ctc.py:453-471 */
};
enum {
    WW_CNN_FP32 = 0,   /* exact fp32, CUDA cores */
    WW_CNN_TENSOR = 1, /* tcgen05 kind::f16; every window inside the calibrated guard band of a decision threshold is
                          re-scored by the fp32 kernel (ww_tc_band_info), so decisions are those of WW_CNN_FP32 */
    WW_CNN_INT8 = 2    /* the DEVICE model: int8 power-of-two twin on tcgen05 kind::i8 fed by int8 rounding + device CMVN
                          (requires WW_CMVN_DEVICE and ww_quantize_weights_i8); logits are out_q * 2^exp_out, exactly */
};

#define WW_CLIP_SAMPLES 16000
#define WW_N_MFCC 13
#define WW_WINDOW_FRAMES 63

/* ---- context ----------------------------------------------------------------------------------- */
int ww_version(void);
/* Create a context on CUDA device `device`.  Fails with WW_ERR_CUDA when no usable GPU is present. */
int ww_create(ww_ctx** out, int device);
void ww_destroy(ww_ctx* ctx);
const char* ww_last_error(const ww_ctx* ctx);

/* Load LightweightKWS weights from HOST arrays in torch state_dict layout
 * (ml_models/src/wakeModel.py:8-27): conv_layers.{0,3,6}.weight [32,13,3] [64,32,3] [128,64,3],
 * classifier.0.weight [64,128], classifier.2.weight [num_classes,64].  All layers are bias-free. */
int ww_load_weights(ww_ctx* ctx, const float* conv1, const float* conv2, const float* conv3,
                    const float* fc1, const float* fc2, int num_classes);

/* frames produced for n_samples: PY 1 + n/256 (torch.stft center=True); ESP (n-320)/256+1 (mfcc.c:448) */
int ww_num_frames(int feat_mode, int n_samples);

/* ---- features: extract_features' inner T.MFCC call / mfcc.c extract_mfcc ---------------------- */
/* pcm: device [n_signals][sig_stride] int16 or fp32 samples, n_samples valid per signal.
 * out: device fp32, [n][13][T] or [n][T][13].  No intermediate is written to device memory. */
int ww_mfcc_batch(ww_ctx* ctx, const void* pcm, int pcm_type, long long n_signals, int n_samples,
                  long long sig_stride, int feat_mode, int layout, float* out, ww_stream_t stream);

/* normalize_mfcc / device CMVN over 63-frame windows [n][13][63] -> [n][13][63] */
int ww_cmvn(ww_ctx* ctx, const float* feats, long long n_windows, int cmvn_mode, float* out,
            ww_stream_t stream);
/* normalize_mfcc(mfcc[n_mfcc, T], method) for rows of ANY length (extract_mfcc.py:47-88): x, out device fp32
 * [n_rows][row_stride] with T valid values per row (in place allowed).  WW_NORM_STANDARD = 'standardization' and
 * 'cmvn' (same arithmetic in the reference: unbiased std, std == 0 -> 1, + 1e-8), WW_NORM_MINMAX = 'minmax'. */
enum { WW_NORM_STANDARD = 0, WW_NORM_MINMAX = 1 };
int ww_normalize_rows(ww_ctx* ctx, const float* x, long long n_rows, int T, long long row_stride, int method,
                      float* out, ww_stream_t stream);

/* ---- model forward: LightweightKWS.forward (ml_models/src/wakeModel.py:29-34) ------------------ */
/* feats[win*win_stride + coef*coef_stride + frame*frame_stride], 63 frames per window.
 * logits: [n_windows][num_classes]; decisions (class 0) may be NULL. */
int ww_cnn_forward(ww_ctx* ctx, const float* feats, long long win_stride, long long coef_stride,
                   long long frame_stride, long long n_windows, int cmvn_mode, int decide_mode,
                   float threshold, int cnn_impl, float* logits, uint8_t* decisions, ww_stream_t stream);

/* ---- int8 power-of-two twin of the model (esp-dl export) ---------- */
/* Quantise the loaded fp32 weights to symmetric per-tensor power-of-two int8 (ml_models/xiaoa.json:5-20).
 * exps[12] = exponents of {input, w1, act1, w2, act2, w3, act3, gap, w_fc1, act_fc1, w_fc2, output}
 * (ml_models/xiaoa.info:3139-3150: -4,-8,-5,-9,-5,-9,-4,-5,-9,-4,-9,-3). */
int ww_quantize_weights_i8(ww_ctx* ctx, const int* exps);
/* x: device int8 [n][13][63] at the input exponent (what detect_task hands to the model,
 * esp_wake_word_detector.cpp:200-220); out: device int8 [n][num_classes] at the output exponent.
 * Integer-exact: reproduces the shipped known-answer vector ml_models/xiaoa.info:3153-3224 (-40). */
int ww_cnn_forward_i8(ww_ctx* ctx, const int8_t* x, long long n_windows, int8_t* out, ww_stream_t stream);
/* Context options.  WW_OPT_I8_IMPL: which kernel ww_cnn_forward_i8 launches -- WW_CNN_TENSOR (default: tcgen05
 * kind::i8, int32 accumulators in TMEM) or WW_CNN_FP32 (here: the CUDA-core integer kernel).  Both are integer-exact
 * and give identical results.
 * WW_OPT_GENERIC_FRONTEND: 1 makes whole-clip launches use the run-time-shaped frontend kernel instead of the
 * instantiation with the 1 s clip shape frozen at compile time (identical results; for A/B timing and tests).
 * WW_OPT_FUSED: how ww_score_clips* run with WW_CNN_TENSOR -- 0 (default): chunked launches (frontend, tcgen05 CNN,
 * exact re-score per chunk); 2: ONE persistent kernel per <= 131 072 clips (frontend pipelines and tcgen05 CNN groups on
 * disjoint SMs of the same launch, features handed over through an L2-resident ring: DRAM traffic = the PCM alone,
 * 0.99 x the algorithmic bytes against 1.21 x, at 26.9 against 30.0 M clips/s on one B200); 1: one kernel from 2048
 * clips on.  Results are bit-identical either way.  WW_OPT_FUSED_CNN_SMS: SMs given to the CNN role (0 = default by CMVN mode).
 * WW_OPT_L2_CHUNK_CLIPS: clips per frontend + CNN launch pair of the chunked tensor path when the features are to stay
 * in L2 (0 = off, default: 131 072-clip chunks through a 429 MB scratch in HBM; 14 208 = 46.5 MB of features that the
 * next chunk overwrites while still in L2 and a whole number of waves for both kernels: DRAM traffic 1.004 x the
 * algorithmic bytes at -1.1 % throughput; one exact re-score launch per 131 072 clips instead of one per chunk).
 * WW_OPT_RESCORE_WINDOW_CLIPS: default hand-over, calls that span more than one 131 072-clip chunk -- clips per exact
 * re-score launch (default 1 048 576: the tcgen05 kernel copies the windows inside the guard band to a compact buffer,
 * sized for the worst case of every window listed, 3.3 KB per clip of the window, allocated at the first such call;
 * the chunks' launch pairs are chained with programmatic dependent launch).  0 = one exact launch per chunk.
 * WW_OPT_CTC_SPLIT: backward pass of ww_ctc_loss_bwd for wide vocabularies (C >= 64, 2S+1 <= 128) -- 1 (default): beta
 * recursion, then fill + patches in one pass over the rows; 2: beta, fill, patches as three launches; 3: beta on a
 * context-owned side stream while the fill streams on the caller's (forked and joined with events), then the patches:
 * measured 2x slower (the recursion's gathers queue behind the fill), kept for A/B; 0: fill, then one recursion kernel
 * that also patches.  Identical results. */
enum { WW_OPT_I8_IMPL = 1, WW_OPT_GENERIC_FRONTEND = 2, WW_OPT_FUSED = 3, WW_OPT_FUSED_CNN_SMS = 4, WW_OPT_L2_CHUNK_CLIPS = 5,
       WW_OPT_RESCORE_WINDOW_CLIPS = 6, WW_OPT_CTC_SPLIT = 7 };
int ww_set_option(ww_ctx* ctx, int option, int value);

/* Guard band of WW_CNN_TENSOR for the weights loaded last.  ww_load_weights runs 4096 calibration windows (noise,
 * hot frames, square waves, full-scale signs, the CMVN extreme point, raw-MFCC-like rows) through both kernels:
 *   |tensor logit - fp32 logit| <= beta * ||window||_F,  beta = min(8 x the largest ratio observed, rigorous bound);
 * the rigorous bound (operator norms of the layers, worst alignment everywhere) is reported for reference, it is
 * 2-3 orders of magnitude above what fp16 operands actually do.  norm_limit: a window with a larger norm could leave
 * the fp16 range in some layer and is always re-scored.  Any pointer may be NULL.  Returns 1 when the tensor kernel
 * is in use for these weights, 0 when WW_CNN_TENSOR falls back to the fp32 kernel for every window (non-finite
 * calibration logits, more than 8 classes), negative on error. */
int ww_tc_band_info(const ww_ctx* ctx, float* beta, float* beta_calibrated, float* beta_rigorous, float* norm_limit);
/* How many windows WW_CNN_TENSOR calls on this context have handed to the fp32 kernel since the counter was last
 * reset (reset != 0 clears it).  Synchronises the device: a reporting call, not a hot-path one. */
long long ww_tc_rescored_total(ww_ctx* ctx, int reset);

/* ---- fused clip scoring: PCM -> MFCC -> CMVN -> CNN -> decision -------------------------------- */
/* pcm: device [n_clips][16000].  By default the features pass from the frontend to the CNN through a context-owned
 * scratch of 131 072 clips (429 MB: it round-trips HBM, 6.5 KB per clip on top of the 32 KB of PCM); per chunk:
 * frontend launch, tcgen05 CNN launch, and one fp32 re-score launch (per call of up to 2^20 clips when the call spans
 * several chunks, WW_OPT_RESCORE_WINDOW_CLIPS).  That is the fastest of three hand-overs; the other two
 * keep the features in L2 (WW_OPT_L2_CHUNK_CLIPS, WW_OPT_FUSED above; measured side by side in DESIGN.md 4.6).
 * One fused call at a time per context: a second host thread gets WW_ERR_BUSY; consecutive calls on different
 * streams are ordered on the device. */
int ww_score_clips(ww_ctx* ctx, const void* pcm, int pcm_type, long long n_clips, int cmvn_mode,
                   int decide_mode, float threshold, int cnn_impl, float* logits, uint8_t* decisions,
                   ww_stream_t stream);
/* same, HOST buffers in and out (chunked, double-buffered H2D / compute / D2H); synchronous */
int ww_score_clips_host(ww_ctx* ctx, const void* pcm_host, int pcm_type, long long n_clips, int cmvn_mode,
                        int decide_mode, float threshold, int cnn_impl, float* logits_host,
                        uint8_t* decisions_host);

/* ---- streaming: sliding 63-frame window at hop 1 frame (esp_wake_word_detector.cpp:10-48,154-263) */
/* pcm: device [n_samples] of ONE stream.  feats_work: device fp32 [13][T], T = ww_num_frames(PY, n).
 * logits: device [T-62][num_classes].  Window w = frames w..w+62, CMVN recomputed per window. */
int ww_stream_score(ww_ctx* ctx, const void* pcm, int pcm_type, long long n_samples, int cmvn_mode,
                    int cnn_impl, float* feats_work, float* logits, ww_stream_t stream);
/* One time segment of a stream (a long stream split over the GPUs of a box, no exchange).
 * pcm: device buffer with samples [first_sample, first_sample + n_samples) of a stream of stream_len samples;
 * computes frames [first_frame, first_frame + n_frames) of the WHOLE stream's frame grid (frame t is centred on
 * sample 256 t; reflect padding only at the true stream ends) into feats_work [13][n_frames] and scores the
 * n_frames - 62 windows made of them into logits.  The buffer must contain every tap of those frames plus one sample
 * before the first (pre-emphasis): 256 first_frame - 161 .. 256 (first_frame + n_frames - 1) + 159, clipped to the
 * stream; WW_ERR_INVALID otherwise.  Stitching the segments ww_b200.shard.stream_segments() describes gives
 * bit for bit what ww_stream_score gives for the whole stream.  first_sample % 8 == 0 keeps the TMA path. */
int ww_stream_score_segment(ww_ctx* ctx, const void* pcm, int pcm_type, long long n_samples, long long first_sample,
                            long long stream_len, long long first_frame, long long n_frames, int cmvn_mode,
                            int cnn_impl, float* feats_work, float* logits, ww_stream_t stream);
/* Host-side hit logic over per-window logits (class 0 column of [n_windows][num_classes]): first score
 * after `warmup` (64) frames, threshold on the logit, `refractory` frames of lock-out then ring reset
 * (esp_wake_word_detector.cpp:38-44,245-258).  Returns the number of hits (<= max_hits written). */
long long ww_stream_events(const float* logits_host, long long n_windows, int num_classes, float threshold_logit,
                           int warmup, int refractory, long long* hits, long long max_hits);

/* ---- streaming sessions: push chunks for many concurrent streams, poll hits ----------- */
/* Push/poll form of the firmware's record_task / detect_task pair and its 63-frame MFCC ring
 * (esp_wake_word_detector.cpp:10-48,52-150,154-263; read_mic_fn / wake_event_callback_t of
 * esp_wake_word_detector.hpp:18-37; intended keep-last-N semantics of main/ring_buffer/ring_buffer.c:57-117).
 * All streams of a session advance together; the scores equal ww_stream_score over the concatenated stream. */
typedef struct ww_session ww_session;
typedef struct {
    int32_t stream;  /* stream index */
    int64_t window;  /* window index (frames window .. window+62) */
    float logit;     /* class-0 logit that crossed the threshold */
} ww_hit;
/* chunk sizes must be multiples of 8 samples; threshold is a logit (ln 4 == sigmoid*100 >= 80);
 * warmup (64) and refractory (313 frames = 5 s) as in ww_stream_events */
int ww_session_open(ww_ctx* ctx, int n_streams, int max_chunk_samples, int cmvn_mode, int cnn_impl,
                    float threshold_logit, int warmup_frames, int refractory_frames, ww_session** out);
/* pcm_host: [n_streams][chunk_samples] int16; computes the newly completed frames, scores the new windows */
int ww_session_write(ww_session* s, const int16_t* pcm_host, int chunk_samples);
/* same, fed with what read_mic delivers: tdm_host [n_streams][12 * chunk_samples] int16 (4 interleaved channels at
 * 48 kHz, esp_wake_word_detector.cpp:92-95); record_task's mix and decimator (cpp:103-121) run on the GPU and yield
 * chunk_samples 16 kHz samples per stream */
int ww_session_write_tdm(ww_session* s, const int16_t* tdm_host, int chunk_samples);
/* drain up to max_hits pending WAKE_WORD_DETECTED events; returns how many were written */
long long ww_session_poll(ww_session* s, ww_hit* hits, long long max_hits);
/* windows scored so far per stream */
long long ww_session_windows(const ww_session* s);
/* logits of the windows scored by the last write: host [n_streams][n][num_classes]; returns n */
long long ww_session_last_logits(const ww_session* s, const float** logits);
void ww_session_close(ww_session* s);

/* ---- CTC best path / keyword (ml_models/test.py:168-217, ml_models/ctc.py:453-471) ------------- */
/* log_probs[t*t_stride + b*b_stride + c]; labels: [B][T] int32 (zero padded), out_len: [B].
 * lengths (valid frames per utterance) and keyword/hits may be NULL.
 * C == 1: the rows are binary logits (keyword posterior of a num_classes=1 model); label 1 iff logit > 0. */
int ww_ctc_greedy(ww_ctx* ctx, const float* log_probs, long long t_stride, long long b_stride, int T, int B,
                  int C, const int32_t* lengths, int decode_mode, int32_t* labels, int32_t* out_len,
                  const int32_t* keyword, int keyword_len, uint8_t* hits, ww_stream_t stream);

/* ---- CTC loss (nn.CTCLoss call shape: ml_models/test.py:89,111-112, ml_models/ctc.py:369,396) -- */
size_t ww_ctc_loss_workspace_bytes(int T, int B, int S);
/* nll: [B] per-sample loss before reduction; zero_infinity as in torch (bit 0).  Bit 1, WW_CTC_BETA_IN_FWD: the caller
 * will also call ww_ctc_loss_bwd with the same flag, the same arguments and unchanged context options -- for wide
 * vocabularies (C >= 64, 2S+1 <= 128), long inputs (T >= 256) and batches that leave the GPU idle (B / 8 CTAs <= half
 * the SMs) the forward call then runs the beta recursion beside the alpha recursion (two
 * latency chains that leave the GPU idle, on a context-owned side stream joined before the call returns the stream)
 * and the backward call is the row-parallel gradient pass alone.  Same results bit for bit. */
#define WW_CTC_BETA_IN_FWD 2
int ww_ctc_loss_fwd(ww_ctx* ctx, const float* log_probs, long long t_stride, long long b_stride, int T, int B,
                    int C, const int32_t* targets, int S, const int32_t* input_lengths,
                    const int32_t* target_lengths, int blank, int zero_infinity, float* nll, void* workspace,
                    ww_stream_t stream);
/* grad[t*gt_stride + b*gb_stride + c] = d(sum_b grad_out[b]*nll[b]) / d(activations), PyTorch convention.
 * workspace: the one ww_ctc_loss_fwd filled (alpha stays intact, so the call may be repeated); the wide-vocabulary
 * path keeps alpha + beta and per-utterance scalars in the rest of it. */
int ww_ctc_loss_bwd(ww_ctx* ctx, const float* log_probs, long long t_stride, long long b_stride, int T, int B,
                    int C, const int32_t* targets, int S, const int32_t* input_lengths,
                    const int32_t* target_lengths, int blank, int zero_infinity, const float* grad_out,
                    void* workspace, float* grad, long long gt_stride, long long gb_stride,
                    ww_stream_t stream);

/* ---- WAV ingestion: main/esp_wav/esp_wav.cpp:8-139, esp_wav.hpp:24-213 ---------------- */
/* The fields of wav::WavHeader after its file constructor, plus the sample count the reference reads. */
typedef struct {
    uint32_t riff_length;     /* esp_wav.cpp:34 */
    uint32_t fmt_length;      /* :66 */
    uint16_t audio_format;    /* :70 (1 = PCM) */
    uint16_t num_channels;    /* :74 */
    uint32_t sample_rate;     /* :78 */
    uint32_t byte_rate;       /* :82 */
    uint16_t block_align;     /* :86 */
    uint16_t bits_per_sample; /* :90 */
    uint32_t data_length;     /* :111, bytes of the data chunk */
    uint32_t raw_data_pos;    /* :133, byte offset of the first sample */
    uint32_t n_samples;       /* min(data_length / 2, max_samples) (:128-132), also bounded by the bytes present */
    int32_t valid;            /* WavHeader::isValid(), esp_wav.hpp:109-118 */
} ww_wav_info;
/* Parse a RIFF/WAVE image held in memory with the reference's rules: fixed tag order RIFF, WAVE, "fmt " (a wrong
 * tag is recorded in `valid` but parsing continues, esp_wav.cpp:29-62), the 16 standard fmt bytes, then chunks are
 * skipped by their size until "data" (:95-121).  WW_ERR_INVALID when the image ends before the data chunk header
 * (the reference logs and returns early).  Host only: no GPU needed. */
int ww_wav_parse(const void* bytes, size_t n_bytes, int max_samples, ww_wav_info* info);
/* Read `n` files with `n_threads` reader threads into pcm_host[n][clip_samples] int16 (pinned or pageable), each
 * truncated / zero-padded to clip_samples as hello_world_main.cpp:196-214 and pad_audio(add_noise_to_pad=False) do.
 * infos[n] and status[n] (WW_OK / WW_ERR_*) may be NULL.  Returns the number of files that failed. */
int ww_wav_load_batch(const char* const* paths, int n, int clip_samples, int n_threads, int16_t* pcm_host,
                      ww_wav_info* infos, int* status);
/* files -> decisions: the reference's entry points walk a directory of WAV files (ml_models/src/extract_mfcc.py:151-176;
 * hello_world_main.cpp:186-278 over /flash/*.wav).  Reads `n` files with `n_threads` reader threads and scores them as a
 * two-buffer pipeline: while the GPU copies in and scores one batch of 16 384 files, the readers fill the other pinned
 * staging buffer with the next.  logits_host [n][num_classes], decisions_host [n] (may be NULL), infos / status as in
 * ww_wav_load_batch; a file that cannot be read is scored as silence and counted.  stats (may be NULL) receives
 * {seconds reading files, seconds blocked on the GPU, total seconds}.  Returns the number of files that failed
 * (>= 0) or a negative WW_ERR_*. */
long long ww_score_wav_files(ww_ctx* ctx, const char* const* paths, long long n, int n_threads, int cmvn_mode,
                             int decide_mode, float threshold, int cnn_impl, float* logits_host,
                             uint8_t* decisions_host, ww_wav_info* infos, int* status, double* stats);
/* Writer side of wav::WavHeader (esp_wav.hpp:41-75,121-213: initialize + write_info_to_file + write_data_to_file +
 * finalize_wav_file): canonical 44-byte header, riff_length = 36 + data bytes. */
int ww_wav_write(const char* path, const int16_t* pcm, size_t n_samples, int num_channels, int sample_rate);

/* ---- front-of-frontend DSP ------------------------------------------------------------- */
/* record_task's 4-channel TDM mix and 48 -> 16 kHz decimator (esp_wake_word_detector.cpp:103-121), bit-exact:
 *   mono = (int16)(((L << 6) + (ref << 5) + (R << 6)) >> 7);  out[i] = (int16)((m[3i] + 2 m[3i+1] + m[3i+2]) >> 2)
 * tdm: device [n_signals][in_stride] int16, 4 interleaved channels at 48 kHz (12 int16 per output sample);
 * pcm_out: device [n_signals][out_stride] int16 mono 16 kHz, n_out samples per signal. */
int ww_tdm_downmix(ww_ctx* ctx, const int16_t* tdm, long long n_signals, long long n_out, long long in_stride,
                   int16_t* pcm_out, long long out_stride, ww_stream_t stream);
/* augment_audio_waveform (ml_models/src/extract_mfcc.py:90-121) on padded clips: audio device fp32 [n][L] ->
 * out device fp32 [n][5][L] = {original, speed 0.8, speed 1.2 (F.interpolate linear, align_corners=False, then
 * pad_audio with zeros / truncation), volume 0.7, volume 1.3 (clamped to [-1, 1])}. */
int ww_augment_waveform(ww_ctx* ctx, const float* audio, long long n, int L, float* out, ww_stream_t stream);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* Test hook for the tensor-core CNN: `dbg_dev` (device, >= 8*31*32 + 8*15*64 + 8*128 + 8*64 floats, or NULL)
 * receives the per-layer activations of the first 8 windows of each launch; *last_rescored receives the
 * number of windows the previous launch handed to the exact fp32 kernel. */
int ww_debug_tc(ww_ctx* ctx, float* dbg_dev, int* last_rescored);
/* Host-only: the C-MFCC (mfcc.c) tables as this machine builds them -- window [320], dense filterbank [40][257], per-filter
 * epsilon bias [40], DCT [40][13] (any pointer may be NULL).  Returns their checksum (0 on failure); the generated
 * kernel code of the ESP mode (csrc/ww_mel_esp.inc, tools/gen_tables.py --esp) is used only when it carries the same one. */
unsigned int ww_debug_esp_tables(float* window320, float* fb_dense, float* bias40, float* dct_40x13);

/* ---- drop-in for main/esp_mfcc/mfcc.h:10-17 ---------------------------------------------------- */
/* Same signature and ownership as the reference's extract_mfcc(): returns a malloc'd
 * float[num_frames * n_mfcc] (frame-major) that the caller releases with ww_free_mfcc(); NULL on bad
 * arguments (mfcc.c:434-437) or when no GPU is available.  Only the reference's fixed parameter set
 * (16000, 320, 256, 512, 40, 13; hello_world_main.cpp:227) is accepted.  The reference's signature has no room for a
 * context, so THIS entry point (and only this one) keeps a process-wide context on device 0 behind a mutex. */
float* ww_extract_mfcc(const float* signal, int signal_len, int sampling_rate, int frame_size, int hop_size,
                       int n_fft, int n_filters, int n_mfcc);
/* the same call with the context explicit: re-entrant across contexts, any device (release with ww_free_mfcc) */
float* ww_extract_mfcc_ctx(ww_ctx* ctx, const float* signal, int signal_len, int sampling_rate, int frame_size,
                           int hop_size, int n_fft, int n_filters, int n_mfcc);
void ww_free_mfcc(float* mfcc);

/* analyze_mfcc_range() of main/esp_mfcc/mfcc.h:16 (mfcc.c:530-553): minimum, maximum and mean over the finite
 * values of a HOST feature array, NaN / Inf skipped.  The reference only logs the line
 * "<label> MFCC Range: min=…, max=…, avg=…, valid=n/size" (ESP_LOGI; ESP_LOGE "No valid values" when none is
 * finite); here the same line goes to stderr when `label` is not NULL and the numbers are returned through `out`
 * (may be NULL).  Returns the number of finite values, or WW_ERR_INVALID for a NULL array or size <= 0 (the
 * reference returns silently).  Host only: no context, no GPU. */
typedef struct {
    float min_val, max_val, avg;
    long long valid, size;
} ww_mfcc_range;
long long ww_analyze_mfcc_range(const float* mfcc_host, long long size, const char* label, ww_mfcc_range* out);

/* ---- host ring buffer: drop-in for main/ring_buffer/ring_buffer.h:17-33 ------------------------------------- */
/* The float ring the firmware stages samples / frames in, with the semantics ring_buffer.c:57-117 intends: a write
 * appends and, once the ring is full, overwrites the oldest values, so the ring always holds the last `buffer_len`
 * values written (a write longer than the ring keeps its last buffer_len values, ring_buffer.c:63-66); a read copies
 * the OLDEST data_len retained values in writing order without consuming them (read_rinbuffer takes a const ring)
 * and fails when fewer are held.  The reference's own code cannot serve as the checker: create_rinbuffer sets
 * end_p = -1 (ring_buffer.c:33), so its first write copies to buffer - 1, and the wrap branch of read_rinbuffer
 * re-reads past start_p (ring_buffer.c:113-114) -- the oracle is oracle/stream.py:RingModel (parity unpinned).
 * Host only, no context, no GPU.  Returns WW_OK or WW_ERR_INVALID (RINBUF_ERROR = -1 in the reference). */
typedef struct ww_ring ww_ring;
int ww_ring_create(ww_ring** out, int buffer_len);                       /* create_rinbuffer */
void ww_ring_delete(ww_ring* r);                                          /* delete_ringbuffer */
int ww_ring_write(ww_ring* r, const float* data, long long data_len);    /* write_rinbuffer */
int ww_ring_read(const ww_ring* r, float* data, int data_len);           /* read_rinbuffer */
int ww_ring_count(const ww_ring* r);                                      /* ringbuf_data_count (static there) */

#ifdef __cplusplus
}
#endif
#endif /* WW_B200_H_ */
