"""WAV ingestion: the reference's `wav::WavHeader` (main/esp_wav/esp_wav.cpp:8-139, esp_wav.hpp:24-213) as a
batched loader that fills pinned int16 buffers for the frontend.

`parse_wav` / `read_wav` follow the reference's parsing rules (fixed RIFF/WAVE/"fmt " order, unknown chunks
skipped by size until "data", at most 16 000 samples); `load_wav_batch` reads many files with native reader
threads straight into one [n, 16000] int16 tensor that `WakeWordScorer.score_host` / `mfcc_batch` consume.
The parser is host code inside libwwb200.so and needs no GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

CLIP_SAMPLES = 16000


def _info_dict(info: L.WavInfo) -> dict:
    return {name: int(getattr(info, name)) for name, _ in L.WavInfo._fields_}


def parse_wav(data: bytes, max_samples: int = CLIP_SAMPLES) -> dict:
    """Header fields of a WAV image in memory (WavHeader's file constructor).  Raises on a truncated image or a
    missing data chunk (the reference logs an error and leaves the header unusable)."""
    lib = L.load_library()
    info = L.WavInfo()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data) if len(data) else (C.c_ubyte * 1)()
    rc = lib.ww_wav_parse(C.cast(buf, C.c_void_p), len(data), int(max_samples), C.byref(info))
    if rc != L.WW_OK:
        raise L.WWError(f"ww_wav_parse failed ({rc}): truncated header or no data chunk")
    return _info_dict(info)


def read_wav(path: str, max_samples: int = CLIP_SAMPLES):
    """-> (int16 numpy [n_samples], info dict).  n_samples = min(data_length / 2, max_samples)."""
    with open(path, "rb") as f:
        data = f.read()
    info = parse_wav(data, max_samples)
    if not info["valid"]:
        raise L.WWError(f"{path}: not a valid PCM WAV (WavHeader::isValid() is false)")
    if info["bits_per_sample"] != 16:
        raise L.WWError(f"{path}: only 16-bit PCM is supported")
    pcm = np.frombuffer(data, dtype="<i2", count=info["n_samples"], offset=info["raw_data_pos"]).copy()
    return pcm, info


def load_wav_batch(paths, clip_samples: int = CLIP_SAMPLES, threads: int | None = None, pinned: bool | None = None,
                   strict: bool = True):
    """Read `paths` into one int16 tensor [n, clip_samples] (truncated / zero padded), in native reader threads.

    Returns (pcm, infos, status): `pcm` is pinned when CUDA is available (so the H2D copy of score_host is
    asynchronous), `infos` a list of header dicts, `status` an int32 numpy array (0 = ok).  With strict=True a
    failed file raises."""
    lib = L.load_library()
    paths = [os.fspath(p) for p in paths]
    n = len(paths)
    if pinned is None:
        pinned = torch.cuda.is_available()
    pcm = torch.zeros((n, clip_samples), dtype=torch.int16, pin_memory=bool(pinned))
    infos = (L.WavInfo * max(n, 1))()
    status = (C.c_int * max(n, 1))()
    arr = (C.c_char_p * max(n, 1))(*[p.encode() for p in paths])
    if threads is None:
        threads = min(32, os.cpu_count() or 1)
    failed = lib.ww_wav_load_batch(arr, n, int(clip_samples), int(threads), C.c_void_p(pcm.data_ptr()), infos, status)
    st = np.array(status[:n], dtype=np.int32)
    if failed < 0:
        raise L.WWError(f"ww_wav_load_batch failed ({failed})")
    if strict and failed:
        bad = [paths[i] for i in range(n) if st[i] != 0]
        raise L.WWError(f"{failed} WAV file(s) could not be loaded: {bad[:5]}")
    return pcm, [_info_dict(infos[i]) for i in range(n)], st


def write_wav(path: str, pcm, sample_rate: int = 16000, num_channels: int = 1) -> None:
    """Canonical 44-byte-header PCM WAV (WavHeader::initialize + write_info_to_file + write_data_to_file +
    finalize_wav_file, esp_wav.hpp:55-213)."""
    lib = L.load_library()
    a = np.ascontiguousarray(np.asarray(pcm, dtype=np.int16))
    rc = lib.ww_wav_write(os.fspath(path).encode(), a.ctypes.data_as(C.c_void_p), a.size, int(num_channels), int(sample_rate))
    if rc != L.WW_OK:
        raise L.WWError(f"ww_wav_write({path}) failed ({rc})")


def score_wav_dir(path, state_dict, device_path=False, threads=None):
    """Score every *.wav of a directory -- the offline check the firmware runs over /flash/*.wav
    (main/hello_world_main.cpp:168-280: load, 63-frame MFCC, model, tally).

    device_path=False: float model, python CMVN, sigmoid(out) > 0.5 (ml_models/main.py:53);
    device_path=True:  int8 rounding + device CMVN + int8 esp-dl model + sigmoid*100 >= 80
                       (esp_wake_word_detector.cpp:128-131,179-258).
    Returns (names, logits [n, C], decisions uint8 [n], number of positives)."""
    from .model import WakeWordScorer

    names = sorted(n for n in os.listdir(path) if n.endswith(".wav"))
    if not names:
        return [], np.zeros((0, 1), np.float32), np.zeros((0,), np.uint8), 0
    pcm, _, _ = load_wav_batch([os.path.join(path, n) for n in names], threads=threads)
    if device_path:
        sc = WakeWordScorer(state_dict, cmvn="device", decision="device", cnn_impl="int8")
    else:
        sc = WakeWordScorer(state_dict, cmvn="python", decision="python", cnn_impl="tensor")
    logits, dec = sc.score_host(pcm)
    return names, logits, dec, int(dec.sum())
