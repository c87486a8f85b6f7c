/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of record_task's TDM mix and decimator,
 * main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:103-121 (20 ms block: 960 TDM frames -> 320 samples;
 * the decimator never crosses a block boundary because 960 = 3 * 320, so a whole stream is processed the same way). */
#include <stdint.h>

void frontdsp_port_tdm_downmix(const int16_t* signal_48k, long long n_out, int16_t* signal_16k) {
    for (long long i = 0; i < n_out; i++) {
        int16_t mono[3];
        for (int f = 0; f < 3; f++) {
            const int16_t* fr = signal_48k + (i * 3 + f) * 4;
            int16_t mic_l = fr[0], aec_ref = fr[1], mic_r = fr[2];                                        /* :104-106 */
            int32_t weighted = (int32_t)((uint32_t)(int32_t)mic_l << 6) + (int32_t)((uint32_t)(int32_t)aec_ref << 5) +
                               (int32_t)((uint32_t)(int32_t)mic_r << 6);                                   /* :108 */
            mono[f] = (int16_t)(weighted >> 7);                                                            /* :109 */
        }
        int32_t w = (int32_t)mono[0] * 1 + (int32_t)mono[1] * 2 + (int32_t)mono[2] * 1;                    /* :116-118 */
        signal_16k[i] = (int16_t)(w >> 2);                                                                 /* :119 */
    }
}
