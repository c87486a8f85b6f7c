"""Drop-in for `ml_models/src/wakeModel.py:4-34` (LightweightKWS) and the engine-level scorer.

`LightweightKWS` keeps the reference's module structure, so a reference `state_dict`
(keys conv_layers.{0,3,6}.weight, classifier.{0,2}.weight) loads unchanged; `forward`
runs in libwwb200.so (inference only -- the reference's training loop is out of scope).
"""
from __future__ import annotations

import ctypes as C
import itertools
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .onnx_reader import load_kws_state_dict

LN4 = math.log(4.0)
_tokens = itertools.count(1)  # unique per weight owner: id() values are recycled after garbage collection


_IMPL = {"fp32": L.CNN_FP32, "tensor": L.CNN_TENSOR, "int8": L.CNN_INT8}
# ml_models/xiaoa.info:3139-3150 (input, w1, act1, w2, act2, w3, act3, gap, w_fc1, act_fc1, w_fc2, output)
XIAOA_EXPONENTS = (-4, -8, -5, -9, -5, -9, -4, -5, -9, -4, -9, -3)


def _impl_code(name):
    try:
        return _IMPL[name]
    except KeyError:
        raise ValueError(f"cnn_impl must be one of {sorted(_IMPL)}") from None


def _push_weights(ctx, sd, owner_key, int8_exponents=None):
    """Make `sd` the context's weight set (no-op when it already is).  int8_exponents: also prepare the int8
    power-of-two twin (cnn_impl='int8')."""
    if getattr(ctx, "weights_owner", None) == owner_key:
        return
    arrs = []
    for k in ("conv_layers.0.weight", "conv_layers.3.weight", "conv_layers.6.weight",
              "classifier.0.weight", "classifier.2.weight"):
        v = sd[k]
        if isinstance(v, torch.Tensor):
            v = v.detach().to("cpu", torch.float32).numpy()
        arrs.append(np.ascontiguousarray(v, dtype=np.float32))
    shapes = [a.shape for a in arrs]
    C_out = shapes[4][0]
    if shapes[:4] != [(32, 13, 3), (64, 32, 3), (128, 64, 3), (64, 128)] or shapes[4] != (C_out, 64):
        raise ValueError(f"not a LightweightKWS state_dict: {shapes}")
    ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs]
    ctx.check(ctx.lib.ww_load_weights(ctx.h, *ptrs, int(C_out)), "ww_load_weights")
    ctx.num_classes = int(C_out)
    if int8_exponents is not None:
        exps = (C.c_int * 12)(*[int(e) for e in int8_exponents])
        ctx.check(ctx.lib.ww_quantize_weights_i8(ctx.h, exps), "ww_quantize_weights_i8")
    ctx.weights_owner = owner_key


class LightweightKWS(nn.Module):
    """3 x [Conv1d(k3,p1,no bias) -> ReLU -> MaxPool1d(2)] 13->32->64->128, GAP,
    Linear(128,64,no bias) -> ReLU -> Linear(64,num_classes,no bias)."""

    def __init__(self, num_classes=3):
        super().__init__()
        self.conv_layers = nn.Sequential(
            nn.Conv1d(13, 32, 3, padding=1, bias=False), nn.ReLU(), nn.MaxPool1d(2),
            nn.Conv1d(32, 64, 3, padding=1, bias=False), nn.ReLU(), nn.MaxPool1d(2),
            nn.Conv1d(64, 128, 3, padding=1, bias=False), nn.ReLU(), nn.MaxPool1d(2),
        )
        self.global_pool = nn.AdaptiveAvgPool1d(1)
        self.classifier = nn.Sequential(
            nn.Linear(128, 64, bias=False), nn.ReLU(), nn.Linear(64, num_classes, bias=False))
        # 'tensor': tcgen05 fp16-operand kernel; every window whose logit lies inside the calibrated guard band of a
        # decision threshold (0 = sigmoid > 0.5, ln 4 = sigmoid*100 >= 80) is re-scored by the exact fp32 kernel, so
        # thresholding the returned logits gives the fp32 path's decisions.  'fp32': the exact kernel for every window.
        self.cnn_impl = "tensor"
        self._ww_token = next(_tokens)

    @classmethod
    def from_onnx(cls, path):
        sd = load_kws_state_dict(path)
        m = cls(num_classes=sd["classifier.2.weight"].shape[0])
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        return m

    def _weights_key(self):
        return ("module", self._ww_token, tuple(p._version for p in self.parameters()),
                tuple(p.data_ptr() for p in self.parameters()))

    def forward(self, x):
        """x: [B, 13, 63] (already normalised) -> logits [B, num_classes]."""
        if x.dim() != 3 or x.shape[1] != 13 or x.shape[2] != 63:
            raise ValueError("LightweightKWS (B200) scores 63-frame windows: expected [B, 13, 63]")
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
            x = x.cuda()
        x = x.to(torch.float32).contiguous()
        ctx = L.get_context(x.device.index)
        _push_weights(ctx, self.state_dict(), self._weights_key())
        B = x.shape[0]
        out = torch.empty((B, ctx.num_classes), dtype=torch.float32, device=x.device)
        impl = _impl_code(self.cnn_impl)
        ctx.check(ctx.lib.ww_cnn_forward(ctx.h, L.ptr(x), 13 * 63, 63, 1, B, L.CMVN_NONE, L.DECIDE_NONE, 0.0,
                                         impl, L.ptr(out), None, L.cur_stream(x.device)), "ww_cnn_forward")
        return out



def forward_int8(state_dict, x_q, exponents=XIAOA_EXPONENTS, device=None, impl="tensor"):
    """int8 power-of-two twin of the model (what esp-dl runs on the device).

    x_q: int8 [B, 13, 63] at the input exponent (-4) -> int8 [B, C] at the output exponent (-3); integer exact.
    impl: 'tensor' (tcgen05 kind::i8, default) or 'cuda' (CUDA-core integer kernel); identical results.
    """
    if not x_q.is_cuda:
        if not torch.cuda.is_available():
            raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
        x_q = x_q.cuda()
    if x_q.dtype != torch.int8 or x_q.dim() != 3 or x_q.shape[1] != 13 or x_q.shape[2] != 63:
        raise ValueError("forward_int8 expects int8 [B, 13, 63]")
    x_q = x_q.contiguous()
    ctx = L.get_context(x_q.device.index if device is None else device)
    key = ("int8", next(_tokens))
    _push_weights(ctx, state_dict, key)
    exps = (C.c_int * 12)(*[int(e) for e in exponents])
    ctx.check(ctx.lib.ww_quantize_weights_i8(ctx.h, exps), "ww_quantize_weights_i8")
    if impl not in ("tensor", "cuda"):
        raise ValueError("impl must be 'tensor' or 'cuda'")
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_I8_IMPL, L.CNN_TENSOR if impl == "tensor" else L.CNN_FP32), "ww_set_option")
    out = torch.empty((x_q.shape[0], ctx.num_classes), dtype=torch.int8, device=x_q.device)
    ctx.check(ctx.lib.ww_cnn_forward_i8(ctx.h, L.ptr(x_q), x_q.shape[0], L.ptr(out), L.cur_stream(x_q.device)),
              "ww_cnn_forward_i8")
    return out


def score_clips_int8(state_dict, pcm, exponents=XIAOA_EXPONENTS, threshold_percent=80.0):
    """The DEVICE decision path end to end on the GPU: PCM -> MFCC -> int8 rounding + device CMVN
    (esp_wake_word_detector.cpp:128-131,179-211) -> model input at exponent -4 (:216-220) -> int8 power-of-two model
    (esp-dl export, ml_models/xiaoa.info) -> sigmoid(out * 2^-3) * 100 >= 80 (:226-228,245).

    pcm: CUDA [B, 16000] int16 / float.  Returns (out_q int8 [B, C] at the output exponent, decisions uint8 [B]).
    Two launches: the frontend and the kind::i8 tensor-core kernel (CMVN fused into its first stage); everything
    after the float features is integer-exact.
    """
    sc = WakeWordScorer(state_dict, cmvn="device", decision="device", cnn_impl="int8", int8_exponents=exponents)
    sc.threshold = float(threshold_percent)
    logits, dec = sc.score(pcm)
    out_q = torch.round(logits * (2.0 ** -int(exponents[11]))).to(torch.int8)   # exact: logits are out_q * 2^exp
    return out_q, dec


class WakeWordScorer:
    """PCM -> logits / decisions: the fused engine call (MFCC + CMVN + CNN + decision).

    decision='python': sigmoid(out) > 0.5 (ml_models/main.py:53);
    decision='device': sigmoid*100 >= 80 with device-style CMVN
    (esp_wake_word_detector.cpp:179-211,226-245).
    """

    def __init__(self, state_dict, device=None, cmvn="python", decision="python", cnn_impl="tensor",
                 int8_exponents=XIAOA_EXPONENTS):
        """cnn_impl: 'tensor' (default: tcgen05 fp16 operands, windows inside the calibrated guard band of the threshold
        re-scored by the fp32 kernel -- decisions are the fp32 path's), 'fp32' (exact, CUDA cores) or
        'int8' (the device model: int8 power-of-two twin on tcgen05 kind::i8; needs cmvn='device')."""
        self.ctx = L.get_context(device)
        self.sd = state_dict
        self._i8 = tuple(int8_exponents) if cnn_impl == "int8" else None
        if cnn_impl == "int8" and cmvn != "device":
            raise ValueError("cnn_impl='int8' is the device path: use cmvn='device'")
        self.cmvn = {"none": L.CMVN_NONE, "python": L.CMVN_PY, "device": L.CMVN_DEVICE}[cmvn]
        if decision == "python":
            self.decide, self.threshold = L.DECIDE_LOGIT, 0.0
        elif decision == "device":
            self.decide, self.threshold = L.DECIDE_DEVICE, 80.0
        else:
            raise ValueError(decision)
        self.cnn_impl = _impl_code(cnn_impl)
        self._key = ("scorer", next(_tokens))
        _push_weights(self.ctx, self.sd, self._key, self._i8)

    @classmethod
    def from_onnx(cls, path, **kw):
        return cls(load_kws_state_dict(path), **kw)

    def _prep(self):
        _push_weights(self.ctx, self.sd, self._key, self._i8)

    def tc_band_info(self):
        """The tensor path's guard band for this scorer's weights (see Context.tc_band_info)."""
        self._prep()
        return self.ctx.tc_band_info()

    def score(self, pcm):
        """pcm: CUDA [B, 16000] int16 or float32 -> (logits [B, C], decisions uint8 [B])."""
        self._prep()
        if not pcm.is_cuda or pcm.dim() != 2 or pcm.shape[1] != 16000:
            raise ValueError("score() expects a CUDA tensor [B, 16000]")
        pcm = pcm.contiguous()
        pcm_type = L.PCM_S16 if pcm.dtype == torch.int16 else L.PCM_F32
        if pcm_type == L.PCM_F32:
            pcm = pcm.to(torch.float32)
        B = pcm.shape[0]
        logits = torch.empty((B, self.ctx.num_classes), dtype=torch.float32, device=pcm.device)
        dec = torch.empty((B,), dtype=torch.uint8, device=pcm.device)
        self.ctx.check(self.ctx.lib.ww_score_clips(self.ctx.h, L.ptr(pcm), pcm_type, B, self.cmvn, self.decide,
                                                   self.threshold, self.cnn_impl, L.ptr(logits), L.ptr(dec),
                                                   L.cur_stream(pcm.device)), "ww_score_clips")
        return logits, dec

    def score_host(self, pcm, logits=None, decisions=None):
        """pcm: HOST [B, 16000] int16/float32 (numpy or CPU tensor, ideally pinned).
        Runs H2D + scoring + D2H inside the C call; returns host numpy arrays."""
        self._prep()
        if isinstance(pcm, torch.Tensor):
            if pcm.is_cuda:
                raise ValueError("score_host() takes host memory")
            arr_ptr, B, dt = pcm.data_ptr(), pcm.shape[0], pcm.dtype
            ok = pcm.is_contiguous() and pcm.shape[1] == 16000
            pcm_type = L.PCM_S16 if dt == torch.int16 else L.PCM_F32
            if dt not in (torch.int16, torch.float32):
                raise ValueError("int16 or float32 PCM expected")
        else:
            pcm = np.ascontiguousarray(pcm)
            arr_ptr, B = pcm.ctypes.data, pcm.shape[0]
            ok = pcm.shape[1] == 16000 and pcm.dtype in (np.int16, np.float32)
            pcm_type = L.PCM_S16 if pcm.dtype == np.int16 else L.PCM_F32
        if not ok:
            raise ValueError("score_host() expects contiguous [B, 16000] int16/float32")
        if logits is None:
            logits = np.empty((B, self.ctx.num_classes), dtype=np.float32)
        if decisions is None:
            decisions = np.empty((B,), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.ww_score_clips_host(
            self.ctx.h, C.c_void_p(arr_ptr), pcm_type, B, self.cmvn, self.decide, self.threshold, self.cnn_impl,
            logits.ctypes.data_as(C.c_void_p), decisions.ctypes.data_as(C.c_void_p)), "ww_score_clips_host")
        return logits, decisions
