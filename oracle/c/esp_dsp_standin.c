/* see shim/esp_dsp.h -- decimation-in-frequency radix-2 (output bit-reversed), then bit reversal */
#include <math.h>
#include "esp_dsp.h"
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
int dsps_fft2r_fc32_ansi(float* d, int n) {
    for (int len = n; len >= 2; len >>= 1) {
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * M_PI * k / len;
                const float wr = (float)cos(ang), wi = (float)sin(ang);
                const int a = i + k, b = i + k + len / 2;
                const float ar = d[2 * a], ai = d[2 * a + 1], br = d[2 * b], bi = d[2 * b + 1];
                d[2 * a] = ar + br; d[2 * a + 1] = ai + bi;
                const float tr = ar - br, ti = ai - bi;
                d[2 * b] = tr * wr - ti * wi; d[2 * b + 1] = tr * wi + ti * wr;
            }
        }
    }
    return 0;
}
int dsps_bit_rev_fc32(float* d, int n) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            float t = d[2 * i]; d[2 * i] = d[2 * j]; d[2 * j] = t;
            t = d[2 * i + 1]; d[2 * i + 1] = d[2 * j + 1]; d[2 * j + 1] = t;
        }
    }
    return 0;
}
int dsps_cplx2reC_fc32(float* d, int n) { (void)d; (void)n; return 0; }
