"""CPU restatement of the reference's WAV reader/writer -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows wav::WavHeader, main/esp_wav/esp_wav.cpp:8-139 (file constructor) and esp_wav.hpp:55-75,109-145
(initialize / isValid / toByteArray).  Pinned by tests/golden/wav_cases.npz, which holds the results of the
reference's OWN esp_wav.cpp (compiled from /root/reference into oracle/_ref/libesp_wav_ref.so by oracle/c/Makefile)
on crafted WAV images; `ref_parse` / `ref_write` call that library directly where it exists.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libesp_wav_ref.so")
FIELDS = ("num_channels", "sample_rate", "bits_per_sample", "data_length", "byte_rate", "block_align", "raw_data_pos",
          "valid")


def parse(data: bytes, max_samples: int = 16000):
    """-> dict of header fields (+ n_samples), or None when the reference would bail out before finding the data
    chunk (esp_wav.cpp returns early: short read or no "data" tag)."""
    pos = 0

    def rd(k):
        nonlocal pos
        if pos + k > len(data):
            return None
        b = data[pos:pos + k]
        pos += k
        return b

    riff = rd(4)                                   # :22-31
    if riff is None:
        return None
    b = rd(4)                                      # :34
    if b is None:
        return None
    riff_length = struct.unpack("<I", b)[0]
    wave = rd(4)                                   # :40-49
    fmt = rd(4) if wave is not None else None      # :52-62
    if wave is None or fmt is None:
        return None
    b = rd(20)                                     # :65-93 fmt_length + the 16 standard bytes (extra fmt bytes are not skipped)
    if b is None:
        return None
    fmt_length, audio_format, num_channels, sample_rate, byte_rate, block_align, bits = struct.unpack("<IHHIIHH", b)
    data_length = None
    while True:                                    # :98-121
        tag = rd(4)
        b = rd(4) if tag is not None else None
        if tag is None or b is None:
            break
        size = struct.unpack("<I", b)[0]
        if tag == b"data":
            data_length = size
            break
        pos = min(pos + size, len(data)) if pos + size > len(data) else pos + size  # fseek past EOF, next fread fails
    if data_length is None:                        # :123-126
        return None
    n = min(data_length // 2, max_samples)         # :128-132
    n = min(n, (len(data) - pos) // 2)
    valid = (riff == b"RIFF" and wave == b"WAVE" and fmt == b"fmt " and audio_format == 1 and num_channels > 0 and
             sample_rate > 0 and bits > 0)         # esp_wav.hpp:109-118
    return dict(riff_length=riff_length, fmt_length=fmt_length, audio_format=audio_format, num_channels=num_channels,
                sample_rate=sample_rate, byte_rate=byte_rate, block_align=block_align, bits_per_sample=bits,
                data_length=data_length, raw_data_pos=pos, n_samples=n, valid=int(valid))


def load_clip(data: bytes, clip_samples: int = 16000):
    """int16 [clip_samples]: samples of the data chunk truncated / zero padded (hello_world_main.cpp:196-214)."""
    info = parse(data, clip_samples)
    if info is None or not info["valid"] or info["bits_per_sample"] != 16:
        return None
    out = np.zeros(clip_samples, dtype=np.int16)
    out[:info["n_samples"]] = np.frombuffer(data, dtype="<i2", count=info["n_samples"], offset=info["raw_data_pos"])
    return out


def header_bytes(n_samples: int, channels: int = 1, sample_rate: int = 16000, bits: int = 16) -> bytes:
    """WavHeader::initialize + toByteArray (esp_wav.hpp:55-75,124-145)."""
    block_align = channels * bits // 8
    data_length = n_samples * 2
    return (b"RIFF" + struct.pack("<I", 36 + data_length) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * block_align, block_align, bits) +
            b"data" + struct.pack("<I", data_length))


def wav_bytes(pcm16, channels: int = 1, sample_rate: int = 16000) -> bytes:
    a = np.ascontiguousarray(np.asarray(pcm16, dtype="<i2"))
    return header_bytes(a.size, channels, sample_rate) + a.tobytes()


# ---- the reference itself (build container only) ------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_parse(path: str):
    lib = C.CDLL(REF_SO)
    f = (C.c_uint32 * 10)()
    lib.esp_wav_ref_parse(os.fspath(path).encode(), f)
    return {k: int(f[i]) for i, k in enumerate(FIELDS)}


def ref_write(path: str, pcm16, channels: int = 1, sample_rate: int = 16000) -> int:
    lib = C.CDLL(REF_SO)
    a = np.ascontiguousarray(np.asarray(pcm16, dtype=np.int16))
    return lib.esp_wav_ref_write(os.fspath(path).encode(), a.ctypes.data_as(C.c_void_p), C.c_uint32(a.size), C.c_uint16(channels),
                                 C.c_uint32(sample_rate))
