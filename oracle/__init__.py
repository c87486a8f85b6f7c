"""CPU oracle for the wake-word hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is product code.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and only as the checker or the timed CPU baseline.
The product (`esp32-wake-word_b200/`) never imports this package and fails
loudly when its CUDA library is missing.

Modules
  mfcc.py    PY-MFCC (torchaudio call sites of ml_models/src/extract_mfcc.py)
             restated twice: through torchaudio itself, and index-level in
             numpy fp64; normalize_mfcc; device CMVN.
  cnn.py     LightweightKWS forward (ml_models/src/wakeModel.py:4-34) and the
             int8 power-of-two fake-quant twin pinned by xiaoa.info's KAT.
  ctc.py     the two greedy decoders (ml_models/test.py:201-217,
             ml_models/ctc.py:453-471), keyword match, CTC loss (alpha/beta in
             numpy + torch.nn.functional.ctc_loss).
  stream.py  sliding 63-frame window scoring and refractory logic
             (main/esp_wake_word_detector/src/esp_wake_word_detector.cpp).
  wav.py     wav::WavHeader reader / writer (main/esp_wav/esp_wav.cpp:8-139,
             esp_wav.hpp), pinned by the reference's own esp_wav.cpp compiled into
             oracle/_ref/libesp_wav_ref.so (tests/golden/wav_cases.npz).
  frontdsp.py  record_task's TDM mix + 48->16 kHz decimator (cpp:103-121, numpy
             and plain C) and augment_audio_waveform (extract_mfcc.py:90-121).
  c/         plain-C restatement of main/esp_mfcc/mfcc.c (C-MFCC, secondary
             mode) and the recipe that compiles the reference's own mfcc.c
             into oracle/_ref/ when /root/reference is present.

Pinning: see each module's header and DESIGN.md section "Oracle pinning".
"""
