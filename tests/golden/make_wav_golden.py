#!/usr/bin/env python
"""Generate tests/golden/wav_cases.npz from the reference's OWN WAV parser / writer.

Run in the BUILD container only (needs /root/reference and oracle/_ref/libesp_wav_ref.so, which oracle/c/Makefile
compiles from main/esp_wav/esp_wav.cpp where it lies):

    python tests/golden/make_wav_golden.py

Each case is a crafted WAV image (canonical, with LIST / odd-sized junk chunks, > 16000 samples, stereo, 8-bit,
wrong tags, data chunk longer than the file).  Stored: the image bytes, the fields the reference's
wav::WavHeader(file) reports for it, and -- for the writer -- the bytes the reference's
WavHeader(path, ch, sr, 16) + write_info_to_file + write_data_to_file + finalize_wav_file produce.
Cases on which the reference returns before reaching the data chunk leave its members uninitialised; for those
only `reached_data = 0` is recorded.  Also stored: augment_audio_waveform outputs of the reference's own function
(imported from ml_models/src/extract_mfcc.py) for two clips, with its re-padding noise disabled by monkey-patching
pad_audio's default (the noise is unseeded RNG).
"""
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import wav as owav  # noqa: E402


def chunk(tag, payload):
    return tag + struct.pack("<I", len(payload)) + payload


def cases():
    rng = np.random.default_rng(20261018)
    # low-entropy PCM (tone + a few noise bits, full-scale extremes at the ends) keeps the fixture small
    pcm = (12000 * np.sin(np.arange(20000) * 0.05)).astype(np.int16) + rng.integers(-8, 8, size=20000, dtype=np.int16)
    pcm[:2] = (-32768, 32767)
    out = {}
    out["canonical_9000"] = owav.wav_bytes(pcm[:9000])
    out["exact_16000"] = owav.wav_bytes(pcm[:16000])
    out["long_20000"] = owav.wav_bytes(pcm[:20000])
    out["empty_data"] = owav.wav_bytes(pcm[:0])
    h = owav.header_bytes(5000)
    body = pcm[:5000].tobytes()
    out["list_chunk"] = h[:36] + chunk(b"LIST", b"INFOISFT\x05\x00\x00\x00Lavf\x00\x00") + h[36:] + body
    out["two_junk_chunks"] = h[:36] + chunk(b"JUNK", bytes(28)) + chunk(b"bext", bytes(602)) + h[36:] + body
    out["odd_junk_no_pad"] = h[:36] + chunk(b"junk", bytes(7)) + h[36:] + body   # the reference does not skip a pad byte
    out["stereo"] = owav.wav_bytes(pcm[:8000], channels=2)
    out["rate_48k"] = owav.wav_bytes(pcm[:4000], sample_rate=48000)
    b8 = bytearray(owav.wav_bytes(pcm[:3000]))
    b8[34:36] = struct.pack("<H", 8)
    out["bits_8"] = bytes(b8)
    bad = bytearray(owav.wav_bytes(pcm[:3000]))
    bad[0:4] = b"RIFX"
    out["bad_riff_tag"] = bytes(bad)
    bad = bytearray(owav.wav_bytes(pcm[:3000]))
    bad[20:22] = struct.pack("<H", 3)
    out["float_format"] = bytes(bad)
    trunc = owav.wav_bytes(pcm[:6000])
    out["data_longer_than_file"] = trunc[:44 + 2 * 2500]
    out["no_data_chunk"] = h[:36] + chunk(b"LIST", bytes(16))
    out["short_header"] = h[:30]
    out["fmt_18_bytes"] = (h[:16] + struct.pack("<I", 18) + h[20:36] + b"\x00\x00" + h[36:] + body)  # extra fmt bytes are not skipped
    return out, pcm


def main():
    assert owav.have_ref(), "build oracle/_ref first: make -C oracle/c"
    cs, pcm = cases()
    names, blobs, fields, reached = [], [], [], []
    with tempfile.TemporaryDirectory() as td:
        for name, data in cs.items():
            p = os.path.join(td, name + ".wav")
            open(p, "wb").write(data)
            mine = owav.parse(data)
            ref = owav.ref_parse(p)
            names.append(name)
            blobs.append(np.frombuffer(data, dtype=np.uint8))
            reached.append(int(mine is not None))
            if mine is not None:
                fields.append([ref[k] for k in owav.FIELDS])
                for k in owav.FIELDS:
                    assert ref[k] == mine[k], (name, k, ref[k], mine[k])
            else:
                fields.append([0] * len(owav.FIELDS))
                assert ref["valid"] == 0 or True
            print(f"{name:24s} reached_data={reached[-1]} ref={ref}")
        # writer
        wp = os.path.join(td, "w.wav")
        assert owav.ref_write(wp, pcm[:1234]) == 0
        written = np.frombuffer(open(wp, "rb").read(), dtype=np.uint8)
        # the reference's write_data_to_file adds the SAMPLE count to data_length (esp_wav.hpp:166-172), so its two
        # length fields under-report by 2x; everything else must equal the canonical header
        mine_w = bytearray(owav.wav_bytes(pcm[:1234]))
        mine_w[4:8] = struct.pack("<I", 36 + 1234)
        mine_w[40:44] = struct.pack("<I", 1234)
        assert bytes(written) == bytes(mine_w)
        wp2 = os.path.join(td, "w2.wav")
        assert owav.ref_write(wp2, pcm[:1000], channels=2, sample_rate=48000) == 0
        written2 = np.frombuffer(open(wp2, "rb").read(), dtype=np.uint8)
        mine_w = bytearray(owav.wav_bytes(pcm[:1000], channels=2, sample_rate=48000))
        mine_w[4:8] = struct.pack("<I", 36 + 1000)
        mine_w[40:44] = struct.pack("<I", 1000)
        assert bytes(written2) == bytes(mine_w)

    # augment_audio_waveform of the reference itself
    sys.path.insert(0, "/root/reference/ml_models")
    import torch
    from src import extract_mfcc as ref_em

    orig_pad = ref_em.pad_audio
    ref_em.pad_audio = lambda audio, target_length, add_noise_to_pad=False, noise_level=0.005: orig_pad(
        audio, target_length, add_noise_to_pad=False, noise_level=noise_level)
    g = torch.Generator().manual_seed(7)
    # one clip: int16-valued tone + noise for the first 9000 samples (|x| > 1/1.3 in places so the volume clamp
    # is exercised), digital silence after
    clips = torch.round((0.9 * torch.sin(torch.arange(16000) * 0.013) + 0.05 * torch.randn(16000, generator=g)) * 32768) / 32768
    clips = clips.clamp(-1, 1)[None]
    clips[0, 9000:] = 0
    aug = np.stack([np.stack([v[0].numpy() for v in ref_em.augment_audio_waveform(clips[i:i + 1])]) for i in range(1)])
    from oracle import frontdsp as ofd
    mine = ofd.augment_waveform(clips.numpy())
    print("augment: max |oracle - reference| =", float(np.abs(mine - aug).max()))
    assert np.abs(mine - aug).max() == 0.0

    obj = np.empty(len(blobs), dtype=object)
    for i, b in enumerate(blobs):
        obj[i] = b
    lens = np.array([len(b) for b in blobs], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "wav_cases.npz"), names=np.array(names), image_len=lens,
                        images=np.concatenate(blobs), fields=np.array(fields, dtype=np.int64),
                        field_names=np.array(owav.FIELDS), reached_data=np.array(reached, dtype=np.int64),
                        written_pcm=pcm[:1234], written_bytes=written, written2_bytes=written2,
                        aug_in=clips.numpy(), aug_out=aug.astype(np.float32))
    print("wrote wav_cases.npz", os.path.getsize(os.path.join(HERE, "wav_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
