"""Drop-in for the reference's feature module `ml_models/src/extract_mfcc.py`.

Same names and argument meaning as the reference (pad_audio :7, add_random_noise :25,
normalize_mfcc :47, augment_audio_waveform :90, extract_features :123), with the
hot arithmetic -- pre-emphasis + T.MFCC (:171-172) and CMVN (:175) -- executed by
libwwb200.so on the GPU.  `extract_features` reads the WAV files with the native batched loader
(ww_b200.wav) and augments on the GPU (ww_b200.frontdsp); the single-clip helpers below keep the
reference's host behaviour.  Batched forms (`mfcc_batch`, `cmvn_batch`) are what large jobs
call directly.
"""
from __future__ import annotations

import os
import wave

import numpy as np
import torch

from . import _lib as L

CLIP_SAMPLES = 16000
N_MFCC = 13
WINDOW_FRAMES = 63


# ---------------------------------------------------------------------------------------------
# host-side glue (same behaviour as the reference helpers)
# ---------------------------------------------------------------------------------------------
def pad_audio(audio, target_length, add_noise_to_pad=True, noise_level=0.005):
    """Pad (noise or zeros) / truncate `audio` [C, N] to `target_length` samples."""
    n = audio.shape[1]
    if n < target_length:
        extra = target_length - n
        if add_noise_to_pad:
            tail = torch.randn(audio.shape[0], extra, dtype=audio.dtype) * noise_level
            return torch.cat([audio, tail.to(audio.device)], dim=1)
        return torch.nn.functional.pad(audio, (0, extra))
    if n > target_length:
        return audio[:, :target_length]
    return audio


def add_random_noise(waveform, noise_level=0.01, snr_range=(5, 20)):
    """White noise at a random SNR (amplitude ratio 10**(snr_db/20)), clamped to [-1, 1]."""
    noise = torch.randn_like(waveform) * noise_level
    lo, hi = snr_range
    snr = 10 ** ((torch.rand(1) * (hi - lo) + lo) / 20)
    p_sig, p_noise = torch.mean(waveform ** 2), torch.mean(noise ** 2)
    if p_noise > 0:
        noise = noise * torch.sqrt(p_sig / (p_noise * snr)).to(noise.device)
    return torch.clamp(waveform + noise, -1.0, 1.0)


def augment_audio_waveform(audio, augment_factor=3):
    """Original + speed 0.8/1.2 (linear interpolation, re-padded to 1 s) + volume 0.7/1.3."""
    out = [audio]
    for speed in (0.8, 1.2):
        size = int(audio.shape[1] * speed)
        y = torch.nn.functional.interpolate(audio.unsqueeze(0), size=size, mode="linear",
                                            align_corners=False).squeeze(0)
        out.append(pad_audio(y, CLIP_SAMPLES))
    for vol in (0.7, 1.3):
        out.append(torch.clamp(audio * vol, -1.0, 1.0))
    return out


def load_wav(path):
    """16-bit PCM WAV -> (float32 [channels, N] in [-1, 1), sample_rate); torchaudio.load semantics."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM is supported")
        n, ch, sr = w.getnframes(), w.getnchannels(), w.getframerate()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch).T
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0), sr


# ---------------------------------------------------------------------------------------------
# GPU calls
# ---------------------------------------------------------------------------------------------
def _as_cuda_2d(x):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dim() == 1:
        x = x[None]
    if x.dim() != 2:
        raise ValueError("expected [B, N] samples")
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
        x = x.cuda(non_blocking=True)
    if x.dtype == torch.int16:
        return x.contiguous(), L.PCM_S16
    return x.to(torch.float32).contiguous(), L.PCM_F32


def mfcc_batch(x, mode="py", layout="coef_major"):
    """Pre-emphasis (0.97) + MFCC of a batch of signals.

    x: [B, N] int16 PCM or float waveform in [-1, 1] (CPU tensors are copied to the current GPU).
    Returns float32 [B, 13, T] (T = 1 + N//256, the reference's T.MFCC output) for mode 'py';
    mode 'esp' follows main/esp_mfcc/mfcc.c (T = (N-320)//256 + 1).  layout 'frame_major' -> [B, T, 13].
    """
    x, pcm_type = _as_cuda_2d(x)
    ctx = L.get_context(x.device.index)
    feat = L.FEAT_PY if mode == "py" else L.FEAT_ESP
    B, N = x.shape
    T = ctx.lib.ww_num_frames(feat, N)
    if T <= 0:
        raise ValueError(f"signal of {N} samples is shorter than one frame")
    lay = L.LAYOUT_COEF_MAJOR if layout == "coef_major" else L.LAYOUT_FRAME_MAJOR
    shape = (B, N_MFCC, T) if lay == L.LAYOUT_COEF_MAJOR else (B, T, N_MFCC)
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    ctx.check(ctx.lib.ww_mfcc_batch(ctx.h, L.ptr(x), pcm_type, B, N, x.stride(0), feat, lay, L.ptr(out),
                                    L.cur_stream(x.device)), "ww_mfcc_batch")
    return out


def cmvn_batch(feats, device_style=False):
    """CMVN of [B, 13, 63] windows on the GPU (python-style unbiased std, or the firmware's)."""
    if feats.dim() != 3 or feats.shape[1] != N_MFCC or feats.shape[2] != WINDOW_FRAMES:
        raise ValueError("cmvn_batch expects [B, 13, 63]")
    x = feats.to(torch.float32).contiguous()
    ctx = L.get_context(x.device.index)
    out = torch.empty_like(x)
    mode = L.CMVN_DEVICE if device_style else L.CMVN_PY
    ctx.check(ctx.lib.ww_cmvn(ctx.h, L.ptr(x), x.shape[0], mode, L.ptr(out), L.cur_stream(x.device)), "ww_cmvn")
    return out


def analyze_mfcc_range(mfcc, label=None):
    """analyze_mfcc_range of main/esp_mfcc/mfcc.h:16 (mfcc.c:530-553): min / max / mean over the finite values of a
    feature array (host work; a CUDA tensor is copied back first).  With `label` the reference's log line goes to
    stderr.  Returns {"min", "max", "avg", "valid", "size"}; `valid` = 0 means no finite value."""
    import ctypes as C

    x = mfcc.detach().cpu().numpy() if isinstance(mfcc, torch.Tensor) else np.asarray(mfcc)
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = L.MfccRange()
    n = L.load_library().ww_analyze_mfcc_range(x.ctypes.data_as(C.c_void_p), x.size,
                                               None if label is None else str(label).encode(), C.byref(out))
    if n < 0:
        raise L.WWError(f"ww_analyze_mfcc_range failed ({n}): empty array")
    return {"min": out.min_val, "max": out.max_val, "avg": out.avg, "valid": int(out.valid), "size": int(out.size)}


def normalize_mfcc(mfcc, method="standardization"):
    """Reference signature (extract_mfcc.py:47): mfcc [13, T] (or [..., T]) -> normalised tensor of the same shape,
    statistics over the last (time) axis.

    'cmvn' and 'standardization' are the same arithmetic in the reference (unbiased std, std == 0 -> 1, + 1e-8);
    'minmax' is (x - min) / (max - min + 1e-8); any other method returns the input, as the reference does.
    All of it runs in libwwb200.so: [.., 13, 63] windows in the CMVN kernel of the hot path, every other shape in
    the generic row kernel (ww_normalize_rows).  A CPU tensor is copied to the GPU and the result copied back (the
    reference is CPU code); without a GPU the call raises -- there is no CPU fallback.
    """
    if method not in ("cmvn", "standardization", "minmax"):
        return mfcc
    if not isinstance(mfcc, torch.Tensor):
        mfcc = torch.as_tensor(mfcc)
    if mfcc.dim() < 1:
        raise ValueError("normalize_mfcc expects at least one axis (time)")
    on_cpu = not mfcc.is_cuda
    if on_cpu and not torch.cuda.is_available():
        raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
    x = (mfcc.cuda() if on_cpu else mfcc).to(torch.float32).contiguous()
    if method != "minmax" and x.dim() >= 2 and x.shape[-1] == WINDOW_FRAMES and x.shape[-2] == N_MFCC:
        out = cmvn_batch(x.reshape(-1, N_MFCC, WINDOW_FRAMES)).reshape(x.shape)
    else:
        T = x.shape[-1]
        rows = x.numel() // T if T else 0
        out = torch.empty_like(x)
        ctx = L.get_context(x.device.index)
        ctx.check(ctx.lib.ww_normalize_rows(ctx.h, L.ptr(x), rows, T, T, L.NORM_MINMAX if method == "minmax"
                                            else L.NORM_STANDARD, L.ptr(out), L.cur_stream(x.device)),
                  "ww_normalize_rows")
    return out.cpu() if on_cpu else out


def _load_clip_batch(paths):
    """Host part of extract_features: the files as one int16 batch [n, 16000] (pinned where CUDA exists) and the
    number of real samples per row.  Multi-channel files keep channel 0 only (torchaudio.load gives [C, N] and the
    reference's `mfcc_transform(...)[0]` keeps the first channel, extract_mfcc.py:154-172)."""
    from . import wav

    pcm, infos, _ = wav.load_wav_batch(paths, clip_samples=CLIP_SAMPLES)
    n_valid = infos.field("n_samples").tolist()
    channels = infos.field("num_channels")
    for k in np.nonzero(channels != 1)[0].tolist():
        c = int(channels[k])
        mono = wav.read_wav(paths[k], max_samples=CLIP_SAMPLES * c)[0][::c]
        pcm[k].zero_()
        pcm[k, :len(mono)] = torch.from_numpy(mono.copy())
        n_valid[k] = len(mono)
    return pcm, n_valid


def extract_features(audio_path="./audio_data/train_data/xiaoa", label=0, is_noise=False,
                     add_noise_to_pad=True, augment_audio=True, normalize_method="cmvn"):
    """Reference signature (extract_mfcc.py:123): walk `audio_path`, return
    (list of [13, 63] feature tensors, list of label tensors).

    The files are read by the native batched WAV loader (ww_b200.wav, esp_wav.cpp's parsing rules) into one
    pinned int16 batch; padding noise, the five augmentation variants (ww_augment_waveform) and all features
    are computed on the GPU, the features of all variants of all files in ONE frontend launch.
    """
    from . import frontdsp

    names = [n for n in os.listdir(audio_path) if n.endswith(".wav")]
    if not names:
        return [], []
    paths = [os.path.join(audio_path, n) for n in names]
    pcm, n_valid = _load_clip_batch(paths)
    audio = pcm.cuda(non_blocking=True).to(torch.float32) / 32768.0          # torchaudio.load normalisation
    if add_noise_to_pad:                                                      # pad_audio(..., noise_level=0.005), :157
        t = torch.arange(CLIP_SAMPLES, device=audio.device)[None, :]
        tail = t >= torch.tensor(n_valid, device=audio.device)[:, None]
        audio = torch.where(tail, torch.randn_like(audio) * 0.005, audio)
    if augment_audio:
        variants = frontdsp.augment_batch(audio)                              # [n, 5, L]
        # the reference re-pads the 0.8-speed variant with pad_audio's DEFAULT (noise 0.005), :107
        size08 = int(CLIP_SAMPLES * 0.8)
        variants[:, 1, size08:] = torch.randn_like(variants[:, 1, size08:]) * 0.005
        clips = variants.reshape(-1, CLIP_SAMPLES)
    else:
        clips = audio
    if is_noise:
        clips = torch.stack([add_random_noise(c[None], noise_level=0.01)[0] for c in clips])
    feats = mfcc_batch(clips)
    feats = normalize_mfcc(feats, method=normalize_method)
    features = [f for f in feats]
    labels = [torch.tensor(label) for _ in features]
    print(f"extracted {len(features)} MFCC features from {len(names)} files (normalisation: {normalize_method})")
    return features, labels
