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
from .model import XIAOA_EXPONENTS, _impl_code, _push_weights, _tokens

WINDOW = 63
LN4 = math.log(4.0)  # sigmoid(x)*100 >= 80  <=>  x >= ln 4


def refractory_frames(seconds=5.0, hop=256, sr=16000):
    """5 s lock-out (vTaskDelay(5000 ms), cpp:248) expressed in frames of the reference hop."""
    return int(math.ceil(seconds * sr / hop))


class StreamScorer:
    def __init__(self, state_dict, device=None, cmvn="device", cnn_impl="tensor", int8_exponents=XIAOA_EXPONENTS):
        """cnn_impl: 'tensor' (default; windows inside the guard band of logit 0 / ln 4 re-scored in fp32), 'fp32', or
        'int8' = the firmware's own model (int8 power-of-two twin, needs cmvn='device')."""
        self.ctx = L.get_context(device)
        self.sd = state_dict
        self.cmvn = {"none": L.CMVN_NONE, "python": L.CMVN_PY, "device": L.CMVN_DEVICE}[cmvn]
        self.cnn_impl = _impl_code(cnn_impl)
        self._i8 = tuple(int8_exponents) if cnn_impl == "int8" else None
        self._key = ("stream", next(_tokens))

    def score(self, pcm):
        """pcm: CUDA [N] int16 / float32 -> (features [13, T], logits [T-62, C])."""
        _push_weights(self.ctx, self.sd, self._key, self._i8)
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

    def score_segment(self, pcm, first_sample, stream_len, first_window, n_windows):
        """One time segment of a long stream (`shard.stream_segments`): pcm = CUDA samples
        [first_sample, first_sample + len(pcm)) of a stream of `stream_len` samples.  Returns
        (features [13, n_windows + 62], logits [n_windows, C]) of windows first_window .. first_window + n_windows - 1
        of the WHOLE stream -- the same bits `score` produces for them (ww_stream_score_segment)."""
        _push_weights(self.ctx, self.sd, self._key, self._i8)
        if not pcm.is_cuda or pcm.dim() != 1:
            raise ValueError("StreamScorer.score_segment expects a 1-D CUDA tensor")
        pcm = pcm.contiguous()
        pcm_type = L.PCM_S16 if pcm.dtype == torch.int16 else L.PCM_F32
        if pcm_type == L.PCM_F32:
            pcm = pcm.to(torch.float32)
        n_frames = int(n_windows) + WINDOW - 1
        feats = torch.empty((13, n_frames), dtype=torch.float32, device=pcm.device)
        logits = torch.empty((int(n_windows), self.ctx.num_classes), dtype=torch.float32, device=pcm.device)
        self.ctx.check(self.ctx.lib.ww_stream_score_segment(
            self.ctx.h, L.ptr(pcm), pcm_type, pcm.numel(), int(first_sample), int(stream_len), int(first_window),
            n_frames, self.cmvn, self.cnn_impl, L.ptr(feats), L.ptr(logits), L.cur_stream(pcm.device)),
            "ww_stream_score_segment")
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


class _Hit(C.Structure):
    _fields_ = [("stream", C.c_int32), ("window", C.c_int64), ("logit", C.c_float)]


class StreamSession:
    """Push/poll scoring of many concurrent streams (the firmware's record/detect task pair as a batch service).

    write(chunk[n_streams, n]) appends n samples (a multiple of 8) to every stream, computes the frames that became
    complete, scores every new 63-frame window and applies the per-stream hit / refractory / ring-reset logic;
    poll() returns the pending (stream, window, logit) events.  Scores equal StreamScorer.score over the
    concatenated stream.
    """

    def __init__(self, state_dict, n_streams, max_chunk_samples=16000, device=None, cmvn="device", cnn_impl="tensor",
                 threshold_logit=LN4, warmup=64, refractory=None, int8_exponents=XIAOA_EXPONENTS):
        self.ctx = L.get_context(device)
        self._key = ("session", next(_tokens))
        _push_weights(self.ctx, state_dict, self._key, tuple(int8_exponents) if cnn_impl == "int8" else None)
        self.n_streams = int(n_streams)
        cm = {"none": L.CMVN_NONE, "python": L.CMVN_PY, "device": L.CMVN_DEVICE}[cmvn]
        impl = _impl_code(cnn_impl)
        if refractory is None:
            refractory = refractory_frames()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.ww_session_open(self.ctx.h, self.n_streams, int(max_chunk_samples), cm, impl,
                                                    float(threshold_logit), int(warmup), int(refractory), C.byref(h)),
                       "ww_session_open")
        self.h = h

    def write(self, chunk):
        """chunk: int16 [n_streams, n] (numpy or CPU tensor). Returns the logits of the newly scored windows
        as a numpy array [n_streams, n_new_windows, C]."""
        if isinstance(chunk, torch.Tensor):
            chunk = chunk.cpu().numpy()
        chunk = np.ascontiguousarray(chunk, dtype=np.int16)
        if chunk.ndim != 2 or chunk.shape[0] != self.n_streams:
            raise ValueError("write() expects int16 [n_streams, n]")
        if getattr(self.ctx, "weights_owner", None) != self._key:
            raise L.WWError("another model's weights were loaded into this context while the session was open")
        self.ctx.check(self.ctx.lib.ww_session_write(self.h, chunk.ctypes.data_as(C.c_void_p), chunk.shape[1]),
                       "ww_session_write")
        p = C.POINTER(C.c_float)()
        n = self.ctx.lib.ww_session_last_logits(self.h, C.byref(p))
        if n <= 0:
            return np.zeros((self.n_streams, 0, self.ctx.num_classes), np.float32)
        return np.ctypeslib.as_array(p, shape=(self.n_streams, n, self.ctx.num_classes)).copy()

    def write_tdm(self, tdm):
        """tdm: int16 [n_streams, 12 * n] -- what the firmware's read_mic delivers (4 interleaved channels at 48 kHz,
        esp_wake_word_detector.cpp:92-95).  The mix + 48 -> 16 kHz decimator of record_task (cpp:103-121) run on the
        GPU; equivalent to write(tdm_downmix(tdm)).  Returns the logits of the newly scored windows."""
        if isinstance(tdm, torch.Tensor):
            tdm = tdm.cpu().numpy()
        tdm = np.ascontiguousarray(tdm, dtype=np.int16)
        if tdm.ndim != 2 or tdm.shape[0] != self.n_streams or tdm.shape[1] % 12:
            raise ValueError("write_tdm() expects int16 [n_streams, 12 * n]")
        if getattr(self.ctx, "weights_owner", None) != self._key:
            raise L.WWError("another model's weights were loaded into this context while the session was open")
        self.ctx.check(self.ctx.lib.ww_session_write_tdm(self.h, tdm.ctypes.data_as(C.c_void_p), tdm.shape[1] // 12),
                       "ww_session_write_tdm")
        p = C.POINTER(C.c_float)()
        n = self.ctx.lib.ww_session_last_logits(self.h, C.byref(p))
        if n <= 0:
            return np.zeros((self.n_streams, 0, self.ctx.num_classes), np.float32)
        return np.ctypeslib.as_array(p, shape=(self.n_streams, n, self.ctx.num_classes)).copy()

    def poll(self, max_hits=4096):
        buf = (_Hit * max_hits)()
        n = self.ctx.lib.ww_session_poll(self.h, buf, max_hits)
        return [(int(buf[i].stream), int(buf[i].window), float(buf[i].logit)) for i in range(n)]

    @property
    def windows(self):
        return int(self.ctx.lib.ww_session_windows(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.ww_session_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RingBuffer:
    """Host float ring, drop-in for main/ring_buffer/ring_buffer.h:17-33 (`create_rinbuffer` / `write_rinbuffer` /
    `read_rinbuffer` / `delete_ringbuffer`): keeps the last `buffer_len` values written; `read(n)` returns the oldest
    n retained values without consuming them (ring_buffer.c:57-117, intended semantics).  Host only."""

    def __init__(self, buffer_len):
        self._lib = L.load_library()
        self._h = C.c_void_p()
        rc = self._lib.ww_ring_create(C.byref(self._h), int(buffer_len))
        if rc != 0:
            raise L.WWError(f"ww_ring_create failed ({rc}): buffer_len must be 1..65535")
        self.buffer_len = int(buffer_len)

    def write(self, data):
        x = np.ascontiguousarray(np.asarray(data, dtype=np.float32).ravel())
        rc = self._lib.ww_ring_write(self._h, x.ctypes.data_as(C.c_void_p), x.size)
        if rc != 0:
            raise L.WWError(f"ww_ring_write failed ({rc}): empty write")

    def __len__(self):
        return int(self._lib.ww_ring_count(self._h))

    def read(self, n):
        out = np.empty(int(n), np.float32)
        rc = self._lib.ww_ring_read(self._h, out.ctypes.data_as(C.c_void_p), int(n))
        if rc != 0:
            raise L.WWError(f"ww_ring_read failed ({rc}): {n} values asked, {len(self)} held")
        return out

    def close(self):
        if self._h:
            self._lib.ww_ring_delete(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
