"""CPU: the WAV and front-DSP oracles against the golden vectors made by the reference itself
(tests/golden/make_wav_golden.py) and against each other; the product's host-only WAV parser / loader / writer
(no GPU needed: pure host code in libwwb200.so) against the oracle and the goldens."""
import os
import struct

import numpy as np
import pytest

from oracle import frontdsp as ofd
from oracle import wav as owav

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cases():
    d = np.load(os.path.join(ROOT, "tests", "golden", "wav_cases.npz"))
    off = np.concatenate([[0], np.cumsum(d["image_len"])])
    imgs = {str(n): bytes(d["images"][off[i]:off[i + 1]]) for i, n in enumerate(d["names"])}
    return d, imgs


@pytest.fixture(scope="module")
def wwlib():
    import __graft_entry__ as ge

    ge.build()
    import ww_b200

    ww_b200.load_library()
    return ww_b200


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's results
# ---------------------------------------------------------------------------------------------------------
def test_oracle_wav_parse_equals_reference(cases):
    d, imgs = cases
    names = [str(k) for k in d["field_names"]]
    assert len(imgs) >= 15
    for i, name in enumerate(d["names"]):
        got = owav.parse(imgs[str(name)])
        if d["reached_data"][i]:
            assert got is not None, name
            for j, k in enumerate(names):
                assert got[k] == int(d["fields"][i][j]), (name, k)
        else:
            assert got is None, name  # the reference bailed out before the data chunk (members uninitialised)


def test_oracle_wav_parse_live_reference(cases, tmp_path):
    """Where oracle/_ref exists (build container) run the reference's esp_wav.cpp itself."""
    if not owav.have_ref():
        pytest.skip("oracle/_ref/libesp_wav_ref.so not built here")
    _, imgs = cases
    for name, data in imgs.items():
        p = tmp_path / (name + ".wav")
        p.write_bytes(data)
        mine = owav.parse(data)
        if mine is None:
            continue
        ref = owav.ref_parse(str(p))
        for k in owav.FIELDS:
            assert ref[k] == mine[k], (name, k)


def test_oracle_tdm_numpy_equals_c_port():
    rng = np.random.default_rng(5)
    x = rng.integers(-32768, 32768, size=12 * 4001, dtype=np.int16)
    # extremes: the (int16_t) cast of the mix wraps when (64 L + 32 ref + 64 R) >> 7 leaves int16
    x[:12] = 32767
    x[12:24] = -32768
    x[24:36] = np.tile(np.array([32767, -32768, 32767, 123], dtype=np.int16), 3)
    a = ofd.tdm_downmix(x)
    b = ofd.tdm_downmix_c(x)
    assert a.dtype == np.int16 and a.shape == (4001,)
    np.testing.assert_array_equal(a, b)
    # all channels at +full scale: weighted = 32767 * 160 >> 7 = 40958 -> wraps to -24578 (cpp:109)
    assert a[0] == np.int16((32767 * 160 >> 7) - 65536) == -24578
    # a 20 ms block is 960 TDM frames -> 320 samples and blocks are independent (960 = 3 * 320)
    np.testing.assert_array_equal(ofd.tdm_downmix(x[:3840]), a[:320])


def test_oracle_augment_equals_reference(cases):
    d, _ = cases
    got = ofd.augment_waveform(d["aug_in"])
    assert got.shape == (1, 5, 16000)
    np.testing.assert_array_equal(got, d["aug_out"])
    assert np.all(got[0, 1, 12800:] == 0) and np.abs(got[0, 3:]).max() <= 1.0


# ---------------------------------------------------------------------------------------------------------
# product (host-only entry points of libwwb200.so) vs oracle / goldens
# ---------------------------------------------------------------------------------------------------------
def test_product_wav_parse(wwlib, cases):
    d, imgs = cases
    for i, name in enumerate(d["names"]):
        data = imgs[str(name)]
        want = owav.parse(data)
        if want is None:
            with pytest.raises(wwlib.WWError):
                wwlib.parse_wav(data)
            continue
        got = wwlib.parse_wav(data)
        for k, v in want.items():
            assert got[k] == v, (name, k)


def test_product_wav_load_batch_and_write(wwlib, cases, tmp_path):
    d, imgs = cases
    good = ["canonical_9000", "exact_16000", "long_20000", "empty_data", "list_chunk", "two_junk_chunks",
            "odd_junk_no_pad", "data_longer_than_file"]
    paths = []
    for name in good:
        p = tmp_path / (name + ".wav")
        p.write_bytes(imgs[name])
        paths.append(str(p))
    pcm, infos, st = wwlib.load_wav_batch(paths, threads=3, pinned=False)
    assert tuple(pcm.shape) == (len(good), 16000) and (st == 0).all()
    for i, name in enumerate(good):
        np.testing.assert_array_equal(pcm[i].numpy(), owav.load_clip(imgs[name]), err_msg=name)
        assert infos[i]["n_samples"] == owav.parse(imgs[name])["n_samples"]
    # failures are reported per file, buffers of failed files are zero
    bad = []
    for name in ["bits_8", "bad_riff_tag", "no_data_chunk", "short_header"]:
        p = tmp_path / (name + ".wav")
        p.write_bytes(imgs[name])
        bad.append(str(p))
    bad.append(str(tmp_path / "does_not_exist.wav"))
    pcm2, _, st2 = wwlib.load_wav_batch(paths[:1] + bad, threads=2, pinned=False, strict=False)
    assert st2[0] == 0 and (st2[1:] != 0).all() and int(pcm2[1:].abs().sum()) == 0
    with pytest.raises(wwlib.WWError):
        wwlib.load_wav_batch(bad, pinned=False)
    # read_wav
    one, info = wwlib.read_wav(paths[2])
    assert one.shape == (16000,) and info["data_length"] == 40000
    # writer: canonical header; equals the reference's own writer output except for the two length fields the
    # reference under-reports (write_data_to_file adds the sample count, esp_wav.hpp:166-172)
    wp = tmp_path / "w.wav"
    wwlib.write_wav(str(wp), d["written_pcm"])
    mine = wp.read_bytes()
    assert mine == owav.wav_bytes(d["written_pcm"])
    ref = bytearray(bytes(d["written_bytes"]))
    n = len(d["written_pcm"])
    assert struct.unpack("<I", ref[40:44])[0] == n  # the quirk: samples, not bytes
    ref[4:8] = struct.pack("<I", 36 + 2 * n)
    ref[40:44] = struct.pack("<I", 2 * n)
    assert mine == bytes(ref)
    back, _ = wwlib.read_wav(str(wp))
    np.testing.assert_array_equal(back, d["written_pcm"])


def test_product_wav_loader_prefix_and_whole_file_paths(wwlib, tmp_path):
    """The batch loader parses the first 4 KiB and reads the samples straight into the batch row; a header that does
    not fit the prefix (big LIST / junk chunks in front of "data") takes the whole-file path.  Both against the
    oracle (and the reference's own parser where it is built), for data chunks at every side of the 4 KiB mark."""
    rng = np.random.default_rng(21)

    def image(junk_sizes, n_samples, truncate=0):
        pcm = rng.integers(-30000, 30000, n_samples).astype(np.int16)
        body = b"".join(b"JUNK" + struct.pack("<I", k) + bytes(rng.integers(0, 256, k, dtype=np.uint8)) for k in junk_sizes)
        fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 1, 1, 16000, 32000, 2, 16)
        data = b"data" + struct.pack("<I", 2 * n_samples) + pcm.tobytes()
        riff = b"WAVE" + fmt + body + data
        img = b"RIFF" + struct.pack("<I", len(riff)) + riff
        return img[: len(img) - truncate] if truncate else img

    cases = {
        "no_junk": image([], 16000),
        "junk_3000": image([3000], 16000),                 # data starts inside the prefix, runs far beyond it
        "junk_4044": image([4044], 16000),                 # "data" tag + size end exactly at byte 4096
        "junk_4045": image([4045], 16000),                 # the size field straddles the prefix end -> whole-file path
        "junk_10000_x2": image([10000, 7000], 12000),      # header far beyond the prefix, short clip (padding)
        "short_700": image([100], 700),                    # whole file inside the prefix
        "truncated": image([2000], 16000, truncate=5001),  # fewer sample bytes than the data chunk announces
        "long": image([5000], 40000),                      # longer than a clip: truncated to 16000
    }
    paths = []
    for name, img in cases.items():
        p = tmp_path / (name + ".wav")
        p.write_bytes(img)
        paths.append(str(p))
    pcm, infos, st = wwlib.load_wav_batch(paths, threads=4, pinned=False)
    assert (st == 0).all()
    for i, (name, img) in enumerate(cases.items()):
        want = owav.parse(img)
        np.testing.assert_array_equal(pcm[i].numpy(), owav.load_clip(img), err_msg=name)
        for k in ("n_samples", "raw_data_pos", "data_length", "riff_length"):
            assert infos[i][k] == want[k], (name, k)
        if owav.have_ref():
            ref = owav.ref_parse(paths[i])
            assert ref["raw_data_pos"] == infos[i]["raw_data_pos"] and ref["data_length"] == infos[i]["data_length"], name


def test_product_wav_loader_single_read_path_rejections(wwlib, tmp_path):
    """Files long enough for the loader's one-read path whose 44 header bytes do not allow it: the row must come out
    as the general path (and the oracle) says -- no stale bytes of the speculative read."""
    rng = np.random.default_rng(22)
    pcm = rng.integers(-30000, 30000, 20000).astype(np.int16)

    def canonical(n_samples, bits=16, riff=b"RIFF", tail=b""):
        h = riff + struct.pack("<I", 36 + 2 * n_samples) + b"WAVE" + \
            struct.pack("<4sIHHIIHH", b"fmt ", 16, 1, 1, 16000, 32000, 2, bits) + b"data" + struct.pack("<I", 2 * n_samples)
        return h + pcm[:n_samples].tobytes() + tail

    trailer = b"LIST" + struct.pack("<I", 30000) + bytes(rng.integers(1, 256, 30000, dtype=np.uint8))
    cases = {
        "ok_exact": (canonical(16000), 0),
        "ok_long_with_trailer": (canonical(20000, tail=trailer), 0),
        "short_data_long_file": (canonical(9000, tail=trailer), 0),      # 9000 samples, then non-audio bytes
        "bits_8_long": (canonical(20000, bits=8), None),
        "bad_riff_long": (canonical(20000, riff=b"RIFX"), None),
    }
    paths = []
    for name, (img, _) in cases.items():
        p = tmp_path / (name + ".wav")
        p.write_bytes(img)
        paths.append(str(p))
    got, infos, st = wwlib.load_wav_batch(paths, threads=2, pinned=False, strict=False)
    for i, (name, (img, want_status)) in enumerate(cases.items()):
        if want_status is None:
            assert st[i] != 0 and int(got[i].abs().sum()) == 0, name
        else:
            assert st[i] == 0, name
            np.testing.assert_array_equal(got[i].numpy(), owav.load_clip(img), err_msg=name)
            assert infos[i]["n_samples"] == owav.parse(img)["n_samples"], name


def test_product_wav_loader_reuses_a_caller_buffer(wwlib, cases, tmp_path):
    import torch

    d, imgs = cases
    names = ["canonical_9000", "exact_16000", "long_20000"]
    paths = []
    for name in names:
        p = tmp_path / (name + ".wav")
        p.write_bytes(imgs[name])
        paths.append(str(p))
    buf = torch.full((8, 16000), 77, dtype=torch.int16)           # stale content must not survive
    pcm, infos, st = wwlib.load_wav_batch(paths, pinned=False, out=buf)
    assert pcm.data_ptr() == buf.data_ptr() and tuple(pcm.shape) == (3, 16000) and (st == 0).all()
    for i, name in enumerate(names):
        np.testing.assert_array_equal(pcm[i].numpy(), owav.load_clip(imgs[name]), err_msg=name)
    assert infos.field("n_samples").tolist() == [9000, 16000, 16000] and len(infos) == 3
    with pytest.raises(ValueError):
        wwlib.load_wav_batch(paths, pinned=False, out=torch.zeros((2, 16000), dtype=torch.int16))
