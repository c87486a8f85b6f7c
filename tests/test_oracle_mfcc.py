"""CPU: pin the PY-MFCC oracle (oracle/mfcc.py) against the reference-generated goldens."""
import math
import os

import numpy as np
import pytest

from oracle import mfcc as om


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_features.npz"))


def test_torchaudio_restatement_matches_reference_extract_features(ref):
    # goldens were produced by the reference's own extract_features (tests/golden/make_golden.py)
    x = om.pcm16_to_float(ref["pcm"])
    got = om.mfcc_torchaudio(x).numpy()
    assert got.shape == ref["mfcc"].shape == (14, 13, 63)
    np.testing.assert_allclose(got, ref["mfcc"], rtol=0, atol=2e-5)


def test_numpy64_restatement_matches_torchaudio(ref):
    x = om.pcm16_to_float(ref["pcm"])
    a = om.mfcc_numpy64(x, tables="torchaudio")
    b = om.mfcc_torchaudio(x, dtype="float64").numpy()
    assert np.abs(a - b).max() < 1e-9  # same tables, fp64 both sides
    c = om.mfcc_numpy64(x, tables="fp64")  # tables regenerated from the published formulas
    assert np.abs(c - ref["mfcc"]).max() < 2e-4


def test_fp32_noise_floor(ref):
    x = om.pcm16_to_float(ref["pcm"])
    a = om.mfcc_torchaudio(x, "float32").numpy()
    b = om.mfcc_torchaudio(x, "float64").numpy()
    assert np.abs(a - b).max() < 2e-4


def test_silent_frame_constant_matches_device_dump(golden_dir):
    z = np.zeros((1, 16000), dtype=np.float32)
    f = om.mfcc_torchaudio(z).numpy()[0]
    c0 = math.sqrt(40.0) * math.log(1e-6)
    assert abs(c0 + 87.377) < 1e-3
    np.testing.assert_allclose(f[0], c0, atol=1e-4)
    np.testing.assert_allclose(f[1:], 0.0, atol=1e-4)
    d = np.load(os.path.join(golden_dir, "device_dumps.npz"))["mfcc_i8"]
    # device dump data1 ends in zero-padded frames whose c0 rounds to -87 (hello_world_main.cpp:67-132)
    assert (d[:, 0, :] == -87).any()


def test_frame_count_and_pad():
    assert om.n_frames(16000) == 63
    assert om.pad_audio(np.ones((1, 10)), 16).shape == (1, 16)
    assert om.pad_audio(np.ones((1, 20)), 16).shape == (1, 16)


def test_normalize_cmvn_matches_reference(ref):
    got = om.normalize_mfcc(ref["mfcc"], "cmvn").numpy()
    np.testing.assert_allclose(got, ref["mfcc_cmvn"], rtol=0, atol=1e-5)  # batched vs per-clip reduction order
    x = ref["mfcc"][0]
    mm = om.normalize_mfcc(x, "minmax").numpy()
    assert mm.min() >= 0 and mm.max() <= 1.0 + 1e-6


def test_cmvn_device_matches_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, "device_dumps.npz"))
    z, q = om.cmvn_device(d["mfcc_i8"].astype(np.float32))
    assert (q == d["cmvn_q"]).all()
    assert np.abs(z).max() <= 127 / 16


def test_lroundf_half_away():
    np.testing.assert_array_equal(om._lroundf(np.array([0.5, -0.5, 1.5, -1.5, 2.4, -2.6], np.float32)),
                                  [1, -1, 2, -2, 2, -3])


def test_synth_is_deterministic():
    a = om.synth_clips_int16(8, seed=1234)
    b = om.synth_clips_int16(4, seed=1234, start_index=4)
    assert (a[4:] == b).all()
    assert (a[3, 9000:] == 0).all() and a[3, :9000].any()


def test_analyze_range_matches_the_references_own_log_lines():
    """oracle.mfcc.analyze_range against what the reference's analyze_mfcc_range (mfcc.c:530-553) logged for the same
    arrays (tests/golden/analyze_range.npz, made by tests/golden/make_golden_range.py from the reference's mfcc.c)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "analyze_range.npz"))
    names = [k[2:] for k in g.files if k.startswith("x_")]
    assert len(names) == 6
    for name in names:
        r = om.analyze_range(g["x_" + name])
        assert om.analyze_range_line(name, r) == str(g["line_" + name]), name
