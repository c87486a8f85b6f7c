/* Pure-C consumer of libwwb200.so (no CUDA headers): what a maintainer's C code does after INTEGRATION.md.
 *   abi_smoke                      -> host-only checks (version, frame counts, loud failure without a GPU)
 *   abi_smoke gpu W.bin PCM.bin N OUT.bin
 *        W.bin  : fp32 conv1[32*13*3] conv2[64*32*3] conv3[128*64*3] fc1[64*128] fc2[64] (torch layout)
 *        PCM.bin: int16 [N][16000]
 *        OUT.bin: fp32 logits[N], uint8 decisions[N], then fp32 mfcc of clip 0 via the mfcc.h drop-in [62*13]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ww_b200.h"

static void* slurp(const char* path, size_t bytes) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    void* p = malloc(bytes);
    if (p && fread(p, 1, bytes, f) != bytes) { free(p); p = NULL; }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    printf("ww_version %d frames(py,16000)=%d frames(esp,16000)=%d\n", ww_version(), ww_num_frames(WW_FEAT_PY, 16000),
           ww_num_frames(WW_FEAT_ESP, 16000));
    if (ww_num_frames(WW_FEAT_PY, 16000) != 63 || ww_num_frames(WW_FEAT_ESP, 16000) != 62) return 2;
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
        ww_ctx* ctx = NULL;
        int rc = ww_create(&ctx, 0);
        printf("ww_create rc=%d\n", rc);
        if (rc == WW_OK) ww_destroy(ctx);
        float x[100] = {0};
        if (ww_extract_mfcc(x, 100, 16000, 320, 256, 512, 40, 13) != NULL) return 3; /* signal_len < frame_size */
        return 0;
    }
    if (argc < 6) return 4;
    const long long n = atoll(argv[4]);
    const size_t nw = 32 * 13 * 3 + 64 * 32 * 3 + 128 * 64 * 3 + 64 * 128 + 64;
    float* w = (float*)slurp(argv[2], nw * sizeof(float));
    int16_t* pcm = (int16_t*)slurp(argv[3], (size_t)n * 16000 * sizeof(int16_t));
    if (!w || !pcm) return 5;
    ww_ctx* ctx = NULL;
    if (ww_create(&ctx, 0) != WW_OK) return 6;
    const float *c1 = w, *c2 = c1 + 32 * 13 * 3, *c3 = c2 + 64 * 32 * 3, *f1 = c3 + 128 * 64 * 3, *f2 = f1 + 64 * 128;
    if (ww_load_weights(ctx, c1, c2, c3, f1, f2, 1) != WW_OK) { fprintf(stderr, "%s\n", ww_last_error(ctx)); return 7; }
    float* logits = (float*)malloc(sizeof(float) * n);
    uint8_t* dec = (uint8_t*)malloc(n);
    int rc = ww_score_clips_host(ctx, pcm, WW_PCM_S16, n, WW_CMVN_PY, WW_DECIDE_LOGIT, 0.0f, WW_CNN_TENSOR, logits, dec);
    if (rc != WW_OK) { fprintf(stderr, "score failed %d: %s\n", rc, ww_last_error(ctx)); return 8; }
    /* error behaviour: bad arguments return WW_ERR_INVALID, like the reference's NULL / ESP_FAIL */
    if (ww_score_clips_host(ctx, NULL, WW_PCM_S16, n, WW_CMVN_PY, WW_DECIDE_LOGIT, 0.0f, WW_CNN_FP32, logits, dec) != WW_ERR_INVALID) return 9;
    float* x = (float*)malloc(sizeof(float) * 16000);
    for (int i = 0; i < 16000; ++i) x[i] = pcm[i] / 32768.0f;
    float* m = ww_extract_mfcc(x, 16000, 16000, 320, 256, 512, 40, 13);
    if (!m) return 10;
    FILE* f = fopen(argv[5], "wb");
    fwrite(logits, sizeof(float), n, f);
    fwrite(dec, 1, n, f);
    fwrite(m, sizeof(float), 62 * 13, f);
    fclose(f);
    ww_free_mfcc(m);
    ww_destroy(ctx);
    printf("scored %lld clips\n", n);
    return 0;
}
