/* C-MFCC oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of main/esp_mfcc/mfcc.c (batch path, extract_mfcc :431-527) of the reference:
 *   pre_emphasis            mfcc.c:66-74     y[0]=x[0]; y[i]=x[i]-0.97*x[i-1]
 *   frame_division          mfcc.c:76-108    frame i = y[256*i .. 256*i+320), no centring, no reflection
 *   apply_window            mfcc.c:110-131   symmetric Hamming 0.53836-(1-0.53836)cos(2*pi*i/319)
 *   compute_power_spectrum  mfcc.c:236-273   zero-pad to 512, FFT, (re^2+im^2)/512 + 1e-12
 *   create_mel_filterbank   mfcc.c:144-234   triangles in FFT-bin index space
 *   apply_mel_filterbank    mfcc.c:275-295   fmaxf(sum, 1e-12)
 *   logf                    mfcc.c:496-498
 *   dct_ii                  mfcc.c:20-64, keep n_mfcc of n_filters (mfcc.c:508-521)
 *
 * PARITY UNPINNED for the FFT stage: the reference calls esp-dsp ^1.7.0 (main/idf_component.yml:19,
 * not vendored) as dsps_fft2r_fc32_ansi + dsps_bit_rev_fc32 + dsps_cplx2reC_fc32 on a zero-imaginary
 * full complex FFT (mfcc.c:259-261); the exact scaling of that call sequence cannot be verified
 * offline and no reference test exercises extract_mfcc.  This port implements the INTENDED math
 * |FFT512(frame)|^2/512 + 1e-12, as does the _ref build (oracle/c/Makefile) through its esp-dsp
 * stand-ins.  Everything except the FFT stage follows the reference's float arithmetic.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static void fft_radix2(float* re, float* im, int n) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * M_PI * k / len;
                const float wr = (float)cos(ang), wi = (float)sin(ang);
                const int a = i + k, b = i + k + len / 2;
                const float xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

static float hz_to_mel(float f) { if (f == 0) f = 1; return 1127.0f * log1pf(f / 700.0f); } /* mfcc.c:133-137 */
static float mel_to_hz(float m) { return 700.0f * (powf(10.0f, m / 2595.0f) - 1.0f); }       /* mfcc.c:139-142 */

/* returns number of frames written, or -1 on invalid arguments (the reference returns NULL) */
int esp_mfcc_port(const float* signal, int signal_len, int sampling_rate, int frame_size, int hop_size, int n_fft,
                  int n_filters, int n_mfcc, float* out /* [num_frames][n_mfcc] */) {
    if (!signal || signal_len < frame_size) return -1;                    /* mfcc.c:434-437 */
    const int num_frames = (signal_len - frame_size) / hop_size + 1;      /* mfcc.c:448 */
    const int nb = n_fft / 2 + 1;
    float* y = (float*)calloc(signal_len, sizeof(float));
    y[0] = signal[0];
    for (int i = 1; i < signal_len; ++i) y[i] = signal[i] - 0.97f * signal[i - 1];
    float* window = (float*)malloc(sizeof(float) * frame_size);
    const float alpha = 0.53836f;
    for (int i = 0; i < frame_size; ++i)
        window[i] = alpha - (1.0f - alpha) * cosf(2.0f * M_PI * i / (frame_size - 1));
    /* filterbank */
    float* fbank = (float*)calloc((size_t)n_filters * nb, sizeof(float));
    {
        const float low_mel = hz_to_mel(0), high_mel = hz_to_mel(sampling_rate / 2);
        int* bins = (int*)malloc(sizeof(int) * (n_filters + 2));
        const float bin_width = (float)sampling_rate / n_fft;
        for (int i = 0; i < n_filters + 2; ++i) {
            const float mel = low_mel + i * (high_mel - low_mel) / (n_filters + 1);
            bins[i] = (int)floorf(mel_to_hz(mel) / bin_width);
        }
        for (int i = 0; i < n_filters; ++i) {
            int left = bins[i], center = bins[i + 1], right = bins[i + 2];
            left = left < 0 ? 0 : left; center = center < 0 ? 0 : center; right = right < 0 ? 0 : right;
            left = left >= nb ? nb - 1 : left; center = center >= nb ? nb - 1 : center; right = right >= nb ? nb - 1 : right;
            if (left >= center) center = left + 1;
            if (center >= right) right = center + 1;
            if (right >= nb) right = nb - 1;
            for (int j = left; j <= center; ++j) if (j >= 0 && j < nb) fbank[i * nb + j] = (float)(j - left) / (center - left);
            for (int j = center; j <= right; ++j) if (j >= 0 && j < nb) fbank[i * nb + j] = (float)(right - j) / (right - center);
        }
        free(bins);
    }
    /* DCT cos table, mfcc.c:36-63 (n_filters = 40 > 32) */
    float* cos_table = (float*)malloc(sizeof(float) * n_filters * n_filters);
    for (int k = 0; k < n_filters; ++k)
        for (int i = 0; i < n_filters; ++i) cos_table[k * n_filters + i] = cosf(M_PI * k * (2 * i + 1) / (2.0f * n_filters));

    float* re = (float*)malloc(sizeof(float) * n_fft);
    float* im = (float*)malloc(sizeof(float) * n_fft);
    float* mel = (float*)malloc(sizeof(float) * n_filters);
    for (int f = 0; f < num_frames; ++f) {
        memset(re, 0, sizeof(float) * n_fft);
        memset(im, 0, sizeof(float) * n_fft);
        for (int j = 0; j < frame_size; ++j) {
            const int s = f * hop_size + j;
            re[j] = (s < signal_len ? y[s] : 0.f) * window[j];
        }
        fft_radix2(re, im, n_fft);
        for (int j = 0; j < n_filters; ++j) {
            float e = 0.f;
            for (int k = 0; k < nb; ++k) {
                const float p = (re[k] * re[k] + im[k] * im[k]) / n_fft + 1e-12f;
                e += p * fbank[j * nb + k];
            }
            mel[j] = logf(fmaxf(e, 1e-12f));
        }
        for (int k = 0; k < n_mfcc && k < n_filters; ++k) {
            float sum = 0.f;
            for (int i = 0; i < n_filters; ++i) sum += mel[i] * cos_table[k * n_filters + i];
            const float scale = (k == 0) ? sqrtf(1.0f / n_filters) : sqrtf(2.0f / n_filters);
            out[f * n_mfcc + k] = scale * sum;
        }
    }
    free(re); free(im); free(mel); free(cos_table); free(fbank); free(window); free(y);
    return num_frames;
}
