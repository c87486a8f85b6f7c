"""CPU: pin the CNN oracle against the reference's shipped KAT and reference-generated goldens."""
import os

import numpy as np

from oracle import cnn, mfcc as om


def test_int8_twin_reproduces_shipped_kat(golden_dir, xiaoa_sd):
    k = np.load(os.path.join(golden_dir, "kat_xiaoa_info.npz"))
    x = k["input_q"].T[None]  # [1, 13, 63]
    out = cnn.forward_int8(x, xiaoa_sd)
    assert int(out[0, 0]) == int(k["output_q"][0]) == -40
    assert out[0, 0] * 2.0 ** int(k["output_exponent"]) == -5.0


def test_fp32_forward_on_kat_input(golden_dir, xiaoa_sd):
    k = np.load(os.path.join(golden_dir, "kat_xiaoa_info.npz"))
    x = k["input_q"].T[None].astype(np.float32) / 16.0
    np.testing.assert_allclose(cnn.forward_torch(x, xiaoa_sd)[0], k["fp32_logit"], atol=1e-5)
    np.testing.assert_allclose(cnn.forward_numpy64(x, xiaoa_sd)[0], k["fp32_logit"], atol=1e-4)
    assert abs(float(k["fp32_logit"][0]) + 4.8535) < 1e-3


def test_device_dumps_logits(golden_dir, xiaoa_sd):
    d = np.load(os.path.join(golden_dir, "device_dumps.npz"))
    z, _ = om.cmvn_device(d["mfcc_i8"].astype(np.float32))
    lg = cnn.forward_torch(z, xiaoa_sd)
    np.testing.assert_allclose(lg, d["logits"], atol=1e-5)
    # data1 -> no wake, data2 -> wake at the firmware's 80 % threshold
    assert list(cnn.decide_device(lg[:, 0])) == [False, True]


def test_forward_matches_reference_model(golden_dir, xiaoa_sd):
    r = np.load(os.path.join(golden_dir, "ref_features.npz"))
    lg = cnn.forward_torch(r["mfcc_cmvn"], xiaoa_sd)
    np.testing.assert_allclose(lg, r["logits"], atol=2e-5)
    lg64 = cnn.forward_numpy64(r["mfcc_cmvn"], xiaoa_sd)
    np.testing.assert_allclose(lg64, r["logits"], atol=1e-4)


def test_decisions():
    x = np.array([-1.0, 0.0, 1e-6, 1.3862, 1.3864, 5.0], dtype=np.float32)
    assert list(cnn.decide_python(x)) == [False, False, True, True, True, True]
    assert list(cnn.decide_device(x)) == [False, False, False, False, True, True]
