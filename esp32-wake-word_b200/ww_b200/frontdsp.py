"""Front-of-frontend DSP on the GPU.

`tdm_downmix`    record_task's 4-channel TDM weighted mix + 48 -> 16 kHz [1, 2, 1]/4 decimator
                 (main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:103-121), bit-exact int16
`augment_batch`  augment_audio_waveform (ml_models/src/extract_mfcc.py:90-121) for a batch of padded clips
"""
from __future__ import annotations

import torch

from . import _lib as L

TDM_CHANNELS = 4
DECIMATION = 3


def _stream():
    return torch.cuda.current_stream().cuda_stream


def tdm_downmix(tdm: torch.Tensor) -> torch.Tensor:
    """tdm: int16 [B, 12 * n] (or [12 * n]) -- frames {MIC-L, AEC ref, MIC-R, unused} at 48 kHz, interleaved.
    Returns int16 [B, n] mono at 16 kHz, exactly the firmware's integer arithmetic."""
    if tdm.dtype != torch.int16:
        raise ValueError("TDM samples must be int16")
    squeeze = tdm.dim() == 1
    x = tdm[None] if squeeze else tdm
    if x.dim() != 2:
        raise ValueError("expected [B, 12 * n] TDM samples")
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
        x = x.cuda(non_blocking=True)
    x = x.contiguous()
    n_out = x.shape[1] // (TDM_CHANNELS * DECIMATION)
    ctx = L.get_context(x.device.index)
    out = torch.empty((x.shape[0], n_out), dtype=torch.int16, device=x.device)
    ctx.check(ctx.lib.ww_tdm_downmix(ctx.h, L.ptr(x), x.shape[0], n_out, x.stride(0), L.ptr(out), out.stride(0) if n_out else 0,
                                     _stream()), "ww_tdm_downmix")
    return out[0] if squeeze else out


def augment_batch(audio: torch.Tensor) -> torch.Tensor:
    """audio: float32 [B, L] padded clips in [-1, 1].  Returns [B, 5, L]: original, speed 0.8, speed 1.2
    (linear interpolation, zero padded / truncated back to L), volume 0.7, volume 1.3 (clamped) -- the
    deterministic part of the reference's augment_audio_waveform (its re-padding noise is unseeded RNG)."""
    if audio.dim() != 2:
        raise ValueError("expected [B, L] clips")
    if not audio.is_cuda:
        if not torch.cuda.is_available():
            raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
        audio = audio.cuda(non_blocking=True)
    x = audio.to(torch.float32).contiguous()
    ctx = L.get_context(x.device.index)
    out = torch.empty((x.shape[0], 5, x.shape[1]), dtype=torch.float32, device=x.device)
    ctx.check(ctx.lib.ww_augment_waveform(ctx.h, L.ptr(x), x.shape[0], x.shape[1], L.ptr(out), _stream()), "ww_augment_waveform")
    return out
