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


class WavInfos:
    """The headers of a loaded batch: behaves like the list of header dicts it replaces (`len`, indexing and
    iteration give dicts) without building a dict per file up front; `field(name)` is the whole column as a numpy
    array (`infos.field("n_samples")`)."""

    def __init__(self, raw, n):
        self._raw = raw                                   # keeps the ctypes array alive
        self._a = np.frombuffer(raw, dtype=np.dtype(L.WavInfo))[:n] if n else np.zeros(0, np.dtype(L.WavInfo))

    def __len__(self):
        return int(self._a.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        row = self._a[i]
        return {name: int(row[name]) for name in self._a.dtype.names}

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def field(self, name):
        return self._a[name]


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
                   strict: bool = True, out: torch.Tensor | None = None):
    """Read `paths` into one int16 tensor [n, clip_samples] (truncated / zero padded), in native reader threads.

    Returns (pcm, infos, status): `pcm` is pinned when CUDA is available (so the H2D copy of score_host is
    asynchronous), `infos` the headers (`WavInfos`: a sequence of header dicts, `.field(name)` for a column), `status` an int32 numpy array (0 = ok).  With strict=True a
    failed file raises.  `out` (int16 CPU tensor with room for [n, clip_samples], contiguous) is filled instead of a
    fresh allocation: a job that walks a large directory in batches reuses one (pinned) buffer and pays the page
    faults of 32 KB per file only once."""
    lib = L.load_library()
    paths = [os.fspath(p) for p in paths]
    n = len(paths)
    if pinned is None:
        pinned = torch.cuda.is_available()
    # every row is written in full by the loader (samples + zero padding, or zeros for a failed file)
    if out is not None:
        if out.dtype != torch.int16 or out.is_cuda or not out.is_contiguous() or out.numel() < n * clip_samples:
            raise ValueError("out must be a contiguous int16 CPU tensor with at least n * clip_samples elements")
        pcm = out.view(-1)[: n * clip_samples].view(n, clip_samples)
    else:
        pcm = torch.empty((n, clip_samples), dtype=torch.int16, pin_memory=bool(pinned))
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
    return pcm, WavInfos(infos, n), st


def write_wav(path: str, pcm, sample_rate: int = 16000, num_channels: int = 1) -> None:
    """Canonical 44-byte-header PCM WAV (WavHeader::initialize + write_info_to_file + write_data_to_file +
    finalize_wav_file, esp_wav.hpp:55-213)."""
    lib = L.load_library()
    a = np.ascontiguousarray(np.asarray(pcm, dtype=np.int16))
    rc = lib.ww_wav_write(os.fspath(path).encode(), a.ctypes.data_as(C.c_void_p), a.size, int(num_channels), int(sample_rate))
    if rc != L.WW_OK:
        raise L.WWError(f"ww_wav_write({path}) failed ({rc})")


def score_wav_files(paths, scorer, threads: int | None = None, strict: bool = True):
    """files -> decisions through `scorer` (a WakeWordScorer) as a two-buffer pipeline inside the library
    (ww_score_wav_files): reader threads fill one pinned batch while the GPU copies in and scores the other.

    Returns (logits [n, C] float32, decisions uint8 [n], infos, status, stats) with stats = {"load_s", "gpu_wait_s",
    "total_s", "files_per_s"}.  A file that cannot be read raises with strict=True, else it is scored as silence."""
    paths = [os.fspath(p) for p in paths]
    n = len(paths)
    scorer._prep()
    ctx = scorer.ctx
    logits = np.empty((n, ctx.num_classes), np.float32)
    dec = np.empty((n,), np.uint8)
    infos = (L.WavInfo * max(n, 1))()
    status = (C.c_int * max(n, 1))()
    arr = (C.c_char_p * max(n, 1))(*[p.encode() for p in paths])
    if threads is None:
        threads = min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    stats = (C.c_double * 3)()
    failed = ctx.lib.ww_score_wav_files(ctx.h, arr, n, int(threads), scorer.cmvn, scorer.decide, scorer.threshold,
                                        scorer.cnn_impl, logits.ctypes.data_as(C.c_void_p),
                                        dec.ctypes.data_as(C.c_void_p), infos, status, stats)
    if failed < 0:
        ctx.check(int(failed), "ww_score_wav_files")
    st = np.array(status[:n], dtype=np.int32)
    if strict and failed:
        bad = [paths[i] for i in range(n) if st[i] != 0]
        raise L.WWError(f"{failed} WAV file(s) could not be loaded: {bad[:5]}")
    out_stats = {"load_s": stats[0], "gpu_wait_s": stats[1], "total_s": stats[2],
                 "files_per_s": n / stats[2] if stats[2] > 0 else float("inf")}
    return logits, dec, WavInfos(infos, n), st, out_stats


def score_wav_dir(path, state_dict, device_path=False, threads=None):
    """Score every *.wav of a directory -- the offline check the firmware runs over /flash/*.wav
    (main/hello_world_main.cpp:168-280: load, 63-frame MFCC, model, tally).

    device_path=False: float model, python CMVN, sigmoid(out) > 0.5 (ml_models/main.py:53);
    device_path=True:  int8 rounding + device CMVN + int8 esp-dl model + sigmoid*100 >= 80
                       (esp_wake_word_detector.cpp:128-131,179-258).
    Reading and scoring overlap (score_wav_files).  Returns (names, logits [n, C], decisions uint8 [n], positives)."""
    from .model import WakeWordScorer

    names = sorted(n for n in os.listdir(path) if n.endswith(".wav"))
    if not names:
        return [], np.zeros((0, 1), np.float32), np.zeros((0,), np.uint8), 0
    if device_path:
        sc = WakeWordScorer(state_dict, cmvn="device", decision="device", cnn_impl="int8")
    else:
        sc = WakeWordScorer(state_dict, cmvn="python", decision="python", cnn_impl="tensor")
    logits, dec, _, _, _ = score_wav_files([os.path.join(path, n) for n in names], sc, threads=threads)
    return names, logits, dec, int(dec.sum())
