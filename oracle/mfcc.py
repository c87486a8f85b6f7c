"""PY-MFCC oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference path restated here: ml_models/src/extract_mfcc.py
  :7-23    pad_audio
  :47-88   normalize_mfcc ('standardization' | 'minmax' | 'cmvn')
  :137-148 T.MFCC(sample_rate=16000, n_mfcc=13, log_mels=True,
           melkwargs={n_fft 512, win_length 320, hop_length 256, n_mels 40,
           window_fn=torch.hamming_window})
  :171     torchaudio.functional.preemphasis(x, coeff=0.97)
  :172     mfcc_transform(preemphasized)[0]
  :175     normalize_mfcc(mfcc, 'cmvn')
and the device twin of CMVN, main/esp_wake_word_detector/src/
esp_wake_word_detector.cpp:179-211.

The arithmetic of T.MFCC lives in third-party torchaudio (not vendored in the
reference, version unpinned there; 2.11.0 in this image).  Two restatements:

  * `mfcc_torchaudio`  -- the reference's own call sequence through the
    installed torchaudio (this IS what the reference executes).
  * `mfcc_numpy64`     -- an index-level fp64 restatement of the published
    algorithm that does not call torchaudio's
    transforms; it only borrows torchaudio's three fp32 constant tables when
    asked to (`tables="torchaudio"`), otherwise regenerates them in fp64.

Pinning: the reference holds no golden float features.  `mfcc_numpy64` is
pinned against `mfcc_torchaudio` (tests/test_oracle_mfcc.py, <= 2e-4 abs),
`mfcc_torchaudio` is pinned against the reference's own
`src/extract_mfcc.py` imported in the build container by
tests/golden/make_golden.py (bit-identical, committed fixture), and the
silent-frame constant c0 = sqrt(40)*ln(1e-6) = -87.377 is checked against the
reference's device dumps (main/hello_world_main.cpp:50-132, value -87).
"""
from __future__ import annotations

import math

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 512
WIN_LENGTH = 320
HOP = 256
N_MELS = 40
N_MFCC = 13
PREEMPH = 0.97
LOG_OFFSET = 1e-6
CLIP_SAMPLES = 16000
N_BINS = N_FFT // 2 + 1


def n_frames(n_samples: int) -> int:
    """torch.stft(center=True): 1 + n_samples // hop."""
    return 1 + n_samples // HOP


# ----------------------------------------------------------------------------
# restatement 1: the reference's call sequence through torchaudio
# ----------------------------------------------------------------------------
_TA_CACHE = {}


def torchaudio_transform(dtype="float32"):
    """The exact transform object extract_mfcc.py:137-148 builds."""
    import torch
    import torchaudio.transforms as T

    key = dtype
    if key not in _TA_CACHE:
        m = T.MFCC(
            sample_rate=SAMPLE_RATE,
            n_mfcc=N_MFCC,
            log_mels=True,
            melkwargs={
                "n_fft": N_FFT,
                "win_length": WIN_LENGTH,
                "hop_length": HOP,
                "n_mels": N_MELS,
                "window_fn": torch.hamming_window,
            },
        )
        if dtype == "float64":
            m = m.double()
        _TA_CACHE[key] = m
    return _TA_CACHE[key]


def mfcc_torchaudio(x, dtype="float32"):
    """x: [B, N] (or [N]) float waveform in [-1, 1] -> [B, 13, 1 + N//256].

    extract_mfcc.py:171-172 -- preemphasis over the whole signal, then T.MFCC.
    """
    import torch
    import torchaudio

    x = torch.as_tensor(x)
    x = x.to(torch.float64 if dtype == "float64" else torch.float32)
    squeeze = x.dim() == 1
    if squeeze:
        x = x[None]
    with torch.no_grad():
        pre = torchaudio.functional.preemphasis(x, coeff=PREEMPH)
        out = torchaudio_transform(dtype)(pre)
    return out[0] if squeeze else out


def pcm16_to_float(pcm):
    """torchaudio.load normalisation of int16 PCM (extract_mfcc.py:154): s / 32768."""
    return np.asarray(pcm, dtype=np.int16).astype(np.float32) / np.float32(32768.0)


def pad_audio(audio, target_length=CLIP_SAMPLES):
    """extract_mfcc.py:7-23 with add_noise_to_pad=False (the noise pad is an
    unseeded RNG draw in the reference and cannot be pinned)."""
    audio = np.asarray(audio)
    n = audio.shape[-1]
    if n < target_length:
        pad = [(0, 0)] * (audio.ndim - 1) + [(0, target_length - n)]
        return np.pad(audio, pad)
    return audio[..., :target_length]


# ----------------------------------------------------------------------------
# restatement 2: index-level fp64
# ----------------------------------------------------------------------------
def hamming_periodic(n=WIN_LENGTH):
    """torch.hamming_window(n) (periodic=True): 0.54 - 0.46 cos(2 pi m / n)."""
    m = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * m / n)


def mel_filterbank_htk(n_bins=N_BINS, f_min=0.0, f_max=8000.0, n_mels=N_MELS):
    """torchaudio.functional.melscale_fbanks(257, 0, 8000, 40, 16000, None, 'htk')."""
    all_freqs = np.linspace(0.0, SAMPLE_RATE // 2, n_bins)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))  # [257, 40]


def dct_ortho(n_mfcc=N_MFCC, n_mels=N_MELS):
    """torchaudio.functional.create_dct(13, 40, 'ortho') -> [40, 13]."""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / n_mels * (n + 0.5) * k)  # [13, 40]
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / n_mels)
    return dct.T


def torchaudio_tables():
    """The three fp32 constant tables torchaudio builds (window, fb, dct)."""
    m = torchaudio_transform("float32")
    w = m.MelSpectrogram.spectrogram.window.numpy().astype(np.float64)
    fb = m.MelSpectrogram.mel_scale.fb.numpy().astype(np.float64)
    dct = m.dct_mat.numpy().astype(np.float64)
    return w, fb, dct


def mfcc_numpy64(x, tables="fp64", return_intermediates=False):
    """Index-level fp64 restatement.  x: [B, N] float -> [B, 13, T]."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    B, N = x.shape
    if tables == "torchaudio":
        w320, fb, dct = torchaudio_tables()
    else:
        w320, fb, dct = hamming_periodic(), mel_filterbank_htk(), dct_ortho()
    # 1. pre-emphasis over the whole signal
    y = x.copy()
    y[:, 1:] -= PREEMPH * x[:, :-1]
    # 2. reflect pad n_fft//2 both sides (torch.stft center=True, pad_mode reflect)
    half = N_FFT // 2
    p = np.pad(y, ((0, 0), (half, half)), mode="reflect")
    # 3. frames, 4. window centred in the 512 frame
    T = 1 + N // HOP
    w = np.zeros(N_FFT)
    off = (N_FFT - WIN_LENGTH) // 2
    w[off:off + WIN_LENGTH] = w320
    idx = HOP * np.arange(T)[:, None] + np.arange(N_FFT)[None, :]
    frames = p[:, idx] * w  # [B, T, 512]
    # 5. rFFT, power (no scaling)
    X = np.fft.rfft(frames, n=N_FFT, axis=-1)
    P = X.real ** 2 + X.imag ** 2  # [B, T, 257]
    # 6-8
    mel = P @ fb
    logmel = np.log(mel + LOG_OFFSET)
    c = logmel @ dct  # [B, T, 13]
    out = np.transpose(c, (0, 2, 1))
    if return_intermediates:
        return out, dict(y=y, frames=frames, power=P, mel=mel, logmel=logmel)
    return out


# ----------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------
def analyze_range(mfcc):
    """analyze_mfcc_range, main/esp_mfcc/mfcc.c:530-553: min / max / mean over the finite values with a FLOAT
    accumulator in array order (the mean of a long array drifts accordingly), and the line the reference logs.
    Pinned by tests/golden/analyze_range.npz (the reference's own function, tests/golden/make_golden_range.py)."""
    x = np.asarray(mfcc, np.float32).ravel()
    ok = np.isfinite(x)
    v = x[ok]
    if v.size == 0:
        return {"min": np.inf, "max": -np.inf, "avg": 0.0, "valid": 0, "size": int(x.size)}
    total = np.float32(0.0)
    # sequential float32 sum (np.sum is pairwise and would hide the drift)
    total = np.add.accumulate(v, dtype=np.float32)[-1]
    return {"min": float(v.min()), "max": float(v.max()), "avg": float(np.float32(total) / np.float32(v.size)),
            "valid": int(v.size), "size": int(x.size)}


def analyze_range_line(label, r):
    """The ESP_LOGI / ESP_LOGE text of mfcc.c:548-551."""
    if r["valid"] == 0:
        return "E %s MFCC: No valid values" % label
    return "%s MFCC Range: min=%.6f, max=%.6f, avg=%.6f, valid=%d/%d" % (label, r["min"], r["max"], r["avg"], r["valid"],
                                                                          r["size"])


def normalize_mfcc(mfcc, method="standardization"):
    """extract_mfcc.py:47-88.  mfcc: [..., 13, T] (torch tensor or ndarray).

    'cmvn' and 'standardization' are the same arithmetic in the reference:
    per-coefficient mean / UNBIASED std over time, std==0 -> 1, eps 1e-8.
    """
    import torch

    t = torch.as_tensor(mfcc)
    if method in ("standardization", "cmvn"):
        mean = t.mean(dim=-1, keepdim=True)
        std = t.std(dim=-1, keepdim=True)
        std = torch.where(std == 0, torch.ones_like(std), std)
        out = (t - mean) / (std + 1e-8)
    elif method == "minmax":
        mn = t.min(dim=-1, keepdim=True)[0]
        mx = t.max(dim=-1, keepdim=True)[0]
        out = (t - mn) / (mx - mn + 1e-8)
    else:
        out = t
    return out


def cmvn_device(mfcc):
    """Device-style CMVN, esp_wake_word_detector.cpp:128-131,179-211.

    mfcc: [..., 13, 63] float features (coef-major here; the device buffer is
    frame-major, the arithmetic is per coefficient either way).
      1. lroundf -> int8 clamp                      (cpp:128-131)
      2. mean, POPULATION std over the 63 frames     (cpp:181-197)
      3. q = clamp(lroundf((x-mean)/(std+1e-8)))     (cpp:200-210)
      4. int8 exponent 0 assigned to the model input at exponent -4
         (cpp:216-220): value saturates at 127/16.
    Returns (z, q): z the float value the model sees, q the int8 CMVN output.
    """
    x = np.asarray(mfcc, dtype=np.float32)
    x8 = np.clip(_lroundf(x), -128, 127).astype(np.float32)
    T = x8.shape[-1]
    mean = x8.sum(axis=-1, keepdims=True, dtype=np.float32) / np.float32(T)
    diff = x8 - mean
    var = (diff * diff).sum(axis=-1, keepdims=True, dtype=np.float32) / np.float32(T)
    std = np.sqrt(var, dtype=np.float32)
    norm = (x8 - mean) / (std + np.float32(1e-8))
    q = np.clip(_lroundf(norm), -128, 127).astype(np.int32)
    q16 = np.clip(q * 16, -128, 127)  # exponent 0 -> exponent -4, int8 saturation
    z = q16.astype(np.float32) / np.float32(16.0)
    return z, q.astype(np.int8)


def _lroundf(x):
    """C lroundf: round half away from zero."""
    x = np.asarray(x, dtype=np.float32)
    return np.where(x >= 0, np.floor(x + np.float32(0.5)), np.ceil(x - np.float32(0.5))).astype(np.int64)


# ----------------------------------------------------------------------------
# synthetic clips -- shared by tests and bench
# ----------------------------------------------------------------------------
def synth_clips_int16(n_clips, seed=1234, n_samples=CLIP_SAMPLES, start_index=0):
    """Four value distributions by clip index mod 4, int16-quantised.

    (0) white N(0, 0.1^2)  (1) uniform [-0.5, 0.5)
    (2) 440 Hz + 3 kHz tones at 0.3 each + N(0, 0.01^2)
    (3) 9000 samples of N(0, 0.1^2) then digital silence
    """
    out = np.empty((n_clips, n_samples), dtype=np.int16)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    for i in range(n_clips):
        idx = start_index + i
        rng = np.random.default_rng([seed, idx])
        k = idx % 4
        if k == 0:
            x = rng.normal(0.0, 0.1, n_samples)
        elif k == 1:
            x = rng.uniform(-0.5, 0.5, n_samples)
        elif k == 2:
            ph = rng.uniform(0, 2 * np.pi, 2)
            x = 0.3 * np.sin(2 * np.pi * 440.0 * t + ph[0]) + 0.3 * np.sin(2 * np.pi * 3000.0 * t + ph[1])
            x = x + rng.normal(0.0, 0.01, n_samples)
        else:
            x = np.zeros(n_samples)
            m = min(9000, n_samples)
            x[:m] = rng.normal(0.0, 0.1, m)
        x = np.clip(x, -1.0, 32767.0 / 32768.0)
        out[i] = np.round(x * 32767.0).astype(np.int16)
    return out
