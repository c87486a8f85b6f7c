"""GPU parity: CMVN, LightweightKWS forward, decisions and the fused clip scorer against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn as ocnn
from oracle import mfcc as om

pytestmark = pytest.mark.gpu


def _model(sd, dev, impl="fp32"):
    """The tests of this file pin the exact fp32 kernel; the tcgen05 default is covered by test_gpu_tc.py."""
    import ww_b200

    m = ww_b200.LightweightKWS(num_classes=sd["classifier.2.weight"].shape[0])
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m.cnn_impl = impl
    return m.to(dev)


def test_known_answer_input_fp32(cuda_device, golden_dir, xiaoa_sd):
    k = np.load(os.path.join(golden_dir, "kat_xiaoa_info.npz"))
    x = torch.from_numpy(k["input_q"].T[None].astype(np.float32) / 16.0).to(cuda_device)
    out = _model(xiaoa_sd, cuda_device)(x).cpu().numpy()
    np.testing.assert_allclose(out[0], k["fp32_logit"], atol=2e-5)


def test_device_dumps(cuda_device, golden_dir, xiaoa_sd):
    """int8 MFCC dumps from the firmware (hello_world_main.cpp:50-132) through device-style CMVN + CNN."""
    import ww_b200
    from ww_b200 import _lib as L

    d = np.load(os.path.join(golden_dir, "device_dumps.npz"))
    feats = torch.from_numpy(d["mfcc_i8"].astype(np.float32)).to(cuda_device)
    z = ww_b200.cmvn_batch(feats, device_style=True).cpu().numpy()
    zo, q = om.cmvn_device(d["mfcc_i8"].astype(np.float32))
    np.testing.assert_array_equal(z, zo)
    ctx = L.get_context(0)
    m = _model(xiaoa_sd, cuda_device)
    m(torch.from_numpy(zo).to(cuda_device))  # pushes the weights
    logits = torch.empty(2, 1, device=cuda_device)
    dec = torch.empty(2, dtype=torch.uint8, device=cuda_device)
    ctx.check(ctx.lib.ww_cnn_forward(ctx.h, L.ptr(feats), 819, 63, 1, 2, L.CMVN_DEVICE, L.DECIDE_DEVICE, 80.0,
                                     L.CNN_FP32, L.ptr(logits), L.ptr(dec), L.cur_stream(cuda_device)), "fwd")
    torch.cuda.synchronize()
    np.testing.assert_allclose(logits.cpu().numpy(), d["logits"], atol=2e-5)
    assert dec.cpu().tolist() == [0, 1]  # data1: no wake, data2: wake (sigmoid 82.6 % >= 80 %)


def test_cmvn_python_style(cuda_device, golden_dir):
    import ww_b200

    r = np.load(os.path.join(golden_dir, "ref_features.npz"))
    f = torch.from_numpy(r["mfcc"]).to(cuda_device)
    got = ww_b200.normalize_mfcc(f, "cmvn").cpu().numpy()
    np.testing.assert_allclose(got, r["mfcc_cmvn"], atol=2e-5)
    one = ww_b200.normalize_mfcc(f[0], "cmvn").cpu().numpy()
    np.testing.assert_allclose(one, r["mfcc_cmvn"][0], atol=2e-5)
    # constant coefficient rows: std == 0 -> 1 (extract_mfcc.py:80)
    c = torch.full((1, 13, 63), 3.0, device=cuda_device)
    assert ww_b200.cmvn_batch(c).abs().max().item() == 0.0


def test_forward_matches_reference_logits(cuda_device, golden_dir, xiaoa_sd):
    r = np.load(os.path.join(golden_dir, "ref_features.npz"))
    out = _model(xiaoa_sd, cuda_device)(torch.from_numpy(r["mfcc_cmvn"]).to(cuda_device)).cpu().numpy()
    np.testing.assert_allclose(out, r["logits"], atol=5e-5)


@pytest.mark.parametrize("num_classes", [1, 3])
def test_forward_random_weights_vs_oracle(cuda_device, num_classes):
    rng = np.random.default_rng(num_classes)
    sd = {
        "conv_layers.0.weight": rng.normal(0, 0.2, (32, 13, 3)).astype(np.float32),
        "conv_layers.3.weight": rng.normal(0, 0.1, (64, 32, 3)).astype(np.float32),
        "conv_layers.6.weight": rng.normal(0, 0.1, (128, 64, 3)).astype(np.float32),
        "classifier.0.weight": rng.normal(0, 0.1, (64, 128)).astype(np.float32),
        "classifier.2.weight": rng.normal(0, 0.2, (num_classes, 64)).astype(np.float32),
    }
    x = rng.normal(size=(300, 13, 63)).astype(np.float32)
    want = ocnn.forward_torch(x, sd)
    got = _model(sd, cuda_device)(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    assert got.shape == (300, num_classes)
    np.testing.assert_allclose(got, want, atol=1e-4, rtol=1e-4)
    want64 = ocnn.forward_numpy64(x[:16], sd)
    np.testing.assert_allclose(got[:16], want64, atol=1e-4, rtol=1e-4)


def test_empty_batch(cuda_device, xiaoa_sd):
    out = _model(xiaoa_sd, cuda_device)(torch.zeros(0, 13, 63, device=cuda_device))
    assert out.shape == (0, 1)


@pytest.mark.parametrize("impl", ["fp32", "tensor"])
@pytest.mark.parametrize("cmvn,decision", [("python", "python"), ("device", "device")])
def test_fused_scorer_vs_oracle_chain(cuda_device, xiaoa_sd, cmvn, decision, impl):
    """PCM -> MFCC -> CMVN -> CNN -> decision.  Logits within 2e-3; decisions EXACT for every clip whose
    oracle margin exceeds the fp32 noise floor (|logit - thr| > 1e-3); the others are counted."""
    import ww_b200

    n = 2048
    pcm = om.synth_clips_int16(n, seed=1234)
    feats = om.mfcc_torchaudio(om.pcm16_to_float(pcm)).numpy()
    if cmvn == "python":
        z = om.normalize_mfcc(feats, "cmvn").numpy()
        thr = 0.0
        decide = ocnn.decide_python
    else:
        z, _ = om.cmvn_device(feats)
        thr = np.log(4.0)
        decide = ocnn.decide_device
    want = ocnn.forward_torch(z, xiaoa_sd)[:, 0]
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cmvn=cmvn, decision=decision, cnn_impl=impl)
    logits, dec = sc.score(torch.from_numpy(pcm).to(cuda_device))
    torch.cuda.synchronize()
    got = logits.cpu().numpy()[:, 0]
    d = dec.cpu().numpy().astype(bool)
    tol = 2e-3 if impl == "fp32" else 1e-2   # fp16 operands away from the threshold (inside the band: fp32 logits)
    if cmvn == "python":
        assert np.abs(got - want).max() < tol
    else:
        # device CMVN rounds features to int8: a feature within float noise of x.5 flips one int8 step;
        # such clips are rare and bounded
        flips = int((np.abs(got - want) > tol).sum())
        print(f"device CMVN, {impl}: {flips}/{n} clips differ by more than {tol} (int8 rounding flips)")
        assert flips < 0.02 * n
    clear = np.abs(want - thr) > 1e-3
    near = int((~clear).sum())
    print(f"{cmvn}: positives {int(decide(want).sum())}/{n}, near-threshold clips excluded: {near}")
    assert near < n * 0.01
    if cmvn == "python":
        assert (d[clear] == decide(want)[clear]).all()
    else:
        agree = (d[clear] == decide(want)[clear])
        assert agree.mean() > 0.995
    # host-buffer path returns exactly the device-buffer results
    lh, dh = sc.score_host(pcm)
    np.testing.assert_array_equal(lh[:, 0], got)
    np.testing.assert_array_equal(dh.astype(bool), d)


def test_fused_scorer_float_input_and_chunking(cuda_device, xiaoa_sd):
    import ww_b200

    n = 16384 + 300  # crosses the L2-sized scratch chunk
    pcm = om.synth_clips_int16(512, seed=77)
    pcm = np.tile(pcm, (n // 512 + 1, 1))[:n]
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)   # default implementation: tcgen05 + guard-band re-score
    l16, d16 = sc.score(torch.from_numpy(pcm).to(cuda_device))
    torch.cuda.synchronize()
    # periodic input -> periodic output, across the chunk boundary too
    assert torch.equal(l16[: n - 512], l16[512:]) and torch.equal(d16[: n - 512], d16[512:])
    lh, dh = sc.score_host(pcm)
    np.testing.assert_array_equal(lh, l16.cpu().numpy())
    xf = torch.from_numpy(om.pcm16_to_float(pcm[:256])).to(cuda_device)
    lf, df = sc.score(xf)
    torch.cuda.synchronize()
    # float PCM takes the float pre-emphasis, int16 PCM the exact integer one (100 x[i] - 97 x[i-1]): features differ by
    # ~1e-4 and the fp16 operands of the default CNN path round them differently
    np.testing.assert_allclose(lf.cpu().numpy(), l16[:256].cpu().numpy(), atol=5e-3)
    assert torch.equal(df, d16[:256]) or (l16[:256, 0].abs().min().item() < 0.1)
