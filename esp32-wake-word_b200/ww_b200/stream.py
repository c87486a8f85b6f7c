"""Streaming sliding-window scoring (batch analogue of the firmware detector).

Stands in for main/esp_wake_word_detector/src/esp_wake_word_detector.cpp: the MFCC ring of
63 frames read oldest -> newest (:10-48), per-window CMVN (:179-211), the model run (:216-223),
and the hit / refractory / ring-reset logic (:245-258).  Features are computed once per frame over
the whole stream; windows advance one frame per step.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .model import _push_weights, _tokens

WINDOW = 63
LN4 = math.log(4.0)  # sigmoid(x)*100 >= 80  <=>  x >= ln 4


def refractory_frames(seconds=5.0, hop=256, sr=16000):
    """5 s lock-out (vTaskDelay(5000 ms), cpp:248) expressed in frames of the reference hop."""
    return int(math.ceil(seconds * sr / hop))


class StreamScorer:
    def __init__(self, state_dict, device=None, cmvn="device", cnn_impl="fp32"):
        self.ctx = L.get_context(device)
        self.sd = state_dict
        self.cmvn = {"none": L.CMVN_NONE, "python": L.CMVN_PY, "device": L.CMVN_DEVICE}[cmvn]
        self.cnn_impl = L.CNN_TENSOR if cnn_impl == "tensor" else L.CNN_FP32
        self._key = ("stream", next(_tokens))

    def score(self, pcm):
        """pcm: CUDA [N] int16 / float32 -> (features [13, T], logits [T-62, C])."""
        _push_weights(self.ctx, self.sd, self._key)
        if not pcm.is_cuda or pcm.dim() != 1:
            raise ValueError("StreamScorer.score expects a 1-D CUDA tensor")
        pcm = pcm.contiguous()
        pcm_type = L.PCM_S16 if pcm.dtype == torch.int16 else L.PCM_F32
        if pcm_type == L.PCM_F32:
            pcm = pcm.to(torch.float32)
        N = pcm.numel()
        T = self.ctx.lib.ww_num_frames(L.FEAT_PY, N)
        if T < WINDOW:
            raise ValueError("stream shorter than one 63-frame window")
        feats = torch.empty((13, T), dtype=torch.float32, device=pcm.device)
        logits = torch.empty((T - WINDOW + 1, self.ctx.num_classes), dtype=torch.float32, device=pcm.device)
        self.ctx.check(self.ctx.lib.ww_stream_score(self.ctx.h, L.ptr(pcm), pcm_type, N, self.cmvn, self.cnn_impl,
                                                    L.ptr(feats), L.ptr(logits), L.cur_stream(pcm.device)),
                       "ww_stream_score")
        return feats, logits


def events(logits, threshold_logit=LN4, warmup=64, refractory=None, max_hits=1 << 20):
    """Window indices that raise WAKE_WORD_DETECTED (host-side, sequential by nature)."""
    lib = L.load_library()
    if refractory is None:
        refractory = refractory_frames()
    if isinstance(logits, torch.Tensor):
        logits = logits.detach().cpu().numpy()
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    if lg.ndim == 1:
        lg = lg[:, None]
    hits = np.empty((max_hits,), dtype=np.int64)
    n = lib.ww_stream_events(lg.ctypes.data_as(C.c_void_p), lg.shape[0], lg.shape[1], float(threshold_logit),
                             int(warmup), int(refractory), hits.ctypes.data_as(C.c_void_p), max_hits)
    if n < 0:
        raise L.WWError(f"ww_stream_events failed ({n})")
    return hits[: min(n, max_hits)].tolist()
