"""Streaming sliding-window oracle -- TEST INFRASTRUCTURE ONLY.

Reference semantics restated: main/esp_wake_word_detector/src/
esp_wake_word_detector.cpp
  :10-48    MFCC ring: keep the last 63 frames x 13 coefficients, read
            oldest -> newest (the intended semantics of main/ring_buffer/
            ring_buffer.c:57-117 as well)
  :38,42,141  the first inference fires when shared_counter == 64, i.e. after
            the 64th frame: the window made of frames 0..62 is never scored
  :179-211  CMVN recomputed per window (population std, int8 rounding)
  :226-228,245  sigmoid*100 >= 80  <=>  logit >= ln 4
  :245-258  after a hit: 5 s refractory, then ring reset (counter back to 0,
            so 64 fresh frames are needed before the next inference)

Features are computed once per frame over the WHOLE stream (hop 256, the
reference hop of ml_models/src/extract_mfcc.py:137-148), windows advance one
frame per step.
"""
from __future__ import annotations

import math

import numpy as np

from . import cnn, mfcc

WINDOW = 63
LN4 = math.log(4.0)


def refractory_frames(seconds=5.0, hop=mfcc.HOP, sr=mfcc.SAMPLE_RATE):
    return int(math.ceil(seconds * sr / hop))


def window_logits(pcm_float, sd, cmvn="python", batch=4096):
    """pcm_float: [N] -> (features [13, T], logits [T-62, C]) on CPU."""
    feats = np.asarray(mfcc.mfcc_torchaudio(pcm_float[None])[0])  # [13, T]
    T = feats.shape[1]
    W = T - WINDOW + 1
    outs = []
    for w0 in range(0, max(W, 0), batch):
        w1 = min(W, w0 + batch)
        idx = np.arange(w0, w1)[:, None] + np.arange(WINDOW)[None, :]
        win = feats[:, idx].transpose(1, 0, 2)  # [w, 13, 63]
        if cmvn == "python":
            z = mfcc.normalize_mfcc(win, "cmvn").numpy()
        elif cmvn == "device":
            z, _ = mfcc.cmvn_device(win)
        else:
            z = win
        outs.append(cnn.forward_torch(z, sd))
    logits = np.concatenate(outs, axis=0) if outs else np.zeros((0, 1), np.float32)
    return feats, logits


def events(logits, threshold_logit=LN4, warmup=64, refractory=None, window=WINDOW):
    """Host-side hit logic over per-window logits (class 0 column).

    Window w covers frames w..w+62 and completes when frame w+62 arrives.
    Returns the list of window indices that raised WAKE_WORD_DETECTED.
    """
    if refractory is None:
        refractory = refractory_frames()
    lg = np.asarray(logits, dtype=np.float32).reshape(len(logits), -1)[:, 0]
    hits = []
    reset_f = 0
    n_frames = len(lg) + window - 1
    f = 0
    while f < n_frames:
        count = f - reset_f + 1
        if count >= warmup:
            w = f - (window - 1)
            if lg[w] >= np.float32(threshold_logit):
                hits.append(w)
                reset_f = f + refractory + 1
                f = reset_f
                continue
        f += 1
    return hits


class RingModel:
    """The float ring of main/ring_buffer/ring_buffer.h:17-33 with the semantics ring_buffer.c:57-117 intends
    (keep the last `buffer_len` values written; read = the oldest `n` retained values, not consumed).

    PARITY UNPINNED: the reference's own implementation cannot produce goldens -- create_rinbuffer sets end_p = -1
    (ring_buffer.c:33), so the first write lands at buffer - 1, and read_rinbuffer's wrap branch copies from
    start_p + remain instead of 0 (ring_buffer.c:113-114).  The scenario of its ring_buffer_test_simple
    (ring_buffer.c:120-200: create 10, write 3, read 3, write 7, read 3) is replayed in tests/test_abi.py."""

    def __init__(self, buffer_len):
        from collections import deque

        self.buffer_len = int(buffer_len)
        self.q = deque(maxlen=self.buffer_len)

    def write(self, data):
        self.q.extend(np.asarray(data, np.float32).tolist())

    def count(self):
        return len(self.q)

    def read(self, n):
        if n <= 0 or n > len(self.q):
            return None                      # RINBUF_ERROR
        return np.array(list(self.q)[:n], np.float32)
