"""CPU restatements of the front-of-frontend DSP -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  tdm_downmix      record_task's TDM mix + decimator, main/esp_wake_word_detector/src/esp_wake_word_detector.cpp:103-121
                   (numpy int32 arithmetic; `tdm_downmix_c` is the plain-C port oracle/c/frontdsp_port.c, gcc-built)
  augment_waveform augment_audio_waveform, ml_models/src/extract_mfcc.py:90-121, with the SAME torch calls the
                   reference makes (F.interpolate linear / clamp) and pad_audio(add_noise_to_pad=False)

Pinning: the integer path has no float freedom -- numpy restatement == plain-C port on random and extreme inputs
(tests/test_oracle_frontdsp.py), including the int16 wrap of the mix for |weighted >> 7| > 32767.  The reference's
.cpp cannot be compiled here (FreeRTOS / esp-dl / I2S dependencies), so the C port restates lines 103-121 operation by operation.
augment_waveform is checked against the reference's own function imported from /root/reference when present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "c", "libfrontdsp_port.so")


def tdm_downmix(tdm):
    """tdm: int16 [..., 12 * n] interleaved 4-channel 48 kHz -> int16 [..., n] mono 16 kHz."""
    x = np.asarray(tdm, dtype=np.int16)
    n = x.shape[-1] // 12
    fr = x[..., :12 * n].reshape(x.shape[:-1] + (3 * n, 4)).astype(np.int32)
    weighted = (fr[..., 0] << 6) + (fr[..., 1] << 5) + (fr[..., 2] << 6)        # cpp:108
    mono = (weighted >> 7).astype(np.int16).astype(np.int32)                      # cpp:109: (int16_t) wraps
    m = mono.reshape(x.shape[:-1] + (n, 3))
    return ((m[..., 0] + 2 * m[..., 1] + m[..., 2]) >> 2).astype(np.int16)        # cpp:116-119


def tdm_downmix_c(tdm):
    if not os.path.exists(PORT_SO):
        subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c"), "libfrontdsp_port.so"], check=True, capture_output=True)
    lib = C.CDLL(PORT_SO)
    lib.frontdsp_port_tdm_downmix.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    x = np.ascontiguousarray(tdm, dtype=np.int16).reshape(-1)
    n = x.size // 12
    out = np.zeros(n, dtype=np.int16)
    lib.frontdsp_port_tdm_downmix(x.ctypes.data, n, out.ctypes.data)
    return out


def augment_waveform(audio):
    """audio: float32 [B, L] -> [B, 5, L]; the reference's torch call sequence (extract_mfcc.py:101-119) with
    zero re-padding (the reference's default re-padding noise is unseeded RNG)."""
    import torch

    x = torch.as_tensor(np.asarray(audio, dtype=np.float32))
    B, L = x.shape
    outs = [x]
    for speed in (0.8, 1.2):
        target = int(L * speed)
        y = torch.nn.functional.interpolate(x.unsqueeze(1), size=target, mode="linear", align_corners=False).squeeze(1)
        y = torch.nn.functional.pad(y, (0, L - target)) if target < L else y[:, :L]
        outs.append(y)
    for vol in (0.7, 1.3):
        outs.append(torch.clamp(x * vol, -1.0, 1.0))
    return torch.stack(outs, dim=1).numpy()
