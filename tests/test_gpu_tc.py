"""GPU parity: the tcgen05 tensor-core CNN path against the oracle, stage by stage and end to end."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cnn as ocnn
from oracle import mfcc as om

pytestmark = pytest.mark.gpu

DBG_FLOATS = 8 * 31 * 32 + 8 * 15 * 64 + 8 * 128 + 8 * 64


def _stages(z, sd):
    """fp32 torch stages for [8,13,63] normalised windows."""
    w = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    x = torch.from_numpy(z)
    a1 = F.max_pool1d(F.relu(F.conv1d(x, w["conv_layers.0.weight"], padding=1)), 2)
    a2 = F.max_pool1d(F.relu(F.conv1d(a1, w["conv_layers.3.weight"], padding=1)), 2)
    a3 = F.max_pool1d(F.relu(F.conv1d(a2, w["conv_layers.6.weight"], padding=1)), 2)
    g = a3.mean(-1)
    h = F.relu(g @ w["classifier.0.weight"].T)
    return a1.numpy(), a2.numpy(), g.numpy(), h.numpy()


def test_layer_by_layer(cuda_device, xiaoa_sd):
    import ww_b200
    from ww_b200 import _lib as L

    rng = np.random.default_rng(0)
    z = rng.normal(size=(8, 13, 63)).astype(np.float32)
    m = ww_b200.LightweightKWS(1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in xiaoa_sd.items()})
    m.cnn_impl = "tensor"
    ctx = L.get_context(0)
    dbg = torch.zeros(DBG_FLOATS, device=cuda_device)
    ctx.check(ctx.lib.ww_debug_tc(ctx.h, L.ptr(dbg), None), "dbg")
    try:
        out = m(torch.from_numpy(z).to(cuda_device))
        torch.cuda.synchronize()
    finally:
        ctx.check(ctx.lib.ww_debug_tc(ctx.h, None, None), "dbg")
    d = dbg.cpu().numpy()
    a1, a2, g, h = _stages(z, xiaoa_sd)
    o = 0
    got1 = d[o:o + 8 * 31 * 32].reshape(8, 31, 32).transpose(0, 2, 1); o += 8 * 31 * 32
    got2 = d[o:o + 8 * 15 * 64].reshape(8, 15, 64).transpose(0, 2, 1); o += 8 * 15 * 64
    gotg = d[o:o + 8 * 128].reshape(8, 128); o += 8 * 128
    goth = d[o:o + 8 * 64].reshape(8, 64)
    for name, got, want in (("conv1", got1, a1), ("conv2", got2, a2), ("gap", gotg, g), ("fc1", goth, h)):
        err = np.abs(got - want).max()
        print(f"{name}: max abs err {err:.3e} (scale {np.abs(want).max():.2f})")
        assert err < 2e-2 * max(1.0, np.abs(want).max()), name
    want = ocnn.forward_torch(z, xiaoa_sd)
    err = np.abs(out.cpu().numpy() - want).max()
    print("logit err", err)
    assert err < 2e-2


@pytest.mark.parametrize("num_classes", [1, 3])
@pytest.mark.parametrize("n", [1, 7, 8, 1000])
def test_forward_vs_oracle(cuda_device, num_classes, n):
    import ww_b200

    rng = np.random.default_rng(num_classes * 100 + n)
    sd = {
        "conv_layers.0.weight": rng.normal(0, 0.2, (32, 13, 3)).astype(np.float32),
        "conv_layers.3.weight": rng.normal(0, 0.1, (64, 32, 3)).astype(np.float32),
        "conv_layers.6.weight": rng.normal(0, 0.1, (128, 64, 3)).astype(np.float32),
        "classifier.0.weight": rng.normal(0, 0.1, (64, 128)).astype(np.float32),
        "classifier.2.weight": rng.normal(0, 0.2, (num_classes, 64)).astype(np.float32),
    }
    x = rng.normal(size=(n, 13, 63)).astype(np.float32)
    m = ww_b200.LightweightKWS(num_classes)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.cnn_impl = "tensor"
    got = m(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    want = ocnn.forward_torch(x, sd)
    scale = max(1.0, np.abs(want).max())
    err = np.abs(got - want).max()
    print(f"C={num_classes} n={n}: err {err:.3e} scale {scale:.2f}")
    assert err < 5e-3 * scale


@pytest.mark.parametrize("cmvn,decision", [("python", "python"), ("device", "device")])
def test_fused_tensor_path_decisions_equal_fp32_path(cuda_device, xiaoa_sd, cmvn, decision):
    """Decisions of the tensor-core path must be IDENTICAL to the fp32 path (borderline windows are re-scored
    by the fp32 kernel); logits agree within the fp16-operand error away from the threshold."""
    import ww_b200
    from ww_b200 import _lib as L

    n = 20000
    pcm = om.synth_clips_int16(2048, seed=4321)
    pcm = np.tile(pcm, (n // 2048 + 1, 1))[:n]
    x = torch.from_numpy(pcm).to(cuda_device)
    ref = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cmvn=cmvn, decision=decision, cnn_impl="fp32")
    l32, d32 = ref.score(x)
    tc = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cmvn=cmvn, decision=decision, cnn_impl="tensor")
    ltc, dtc = tc.score(x)
    torch.cuda.synchronize()
    ctx = L.get_context(0)
    k = C.c_int(0)
    ctx.check(ctx.lib.ww_debug_tc(ctx.h, None, C.byref(k)), "dbg")
    err = (ltc - l32).abs().max().item()
    print(f"{cmvn}: max |logit_tc - logit_fp32| = {err:.3e}; re-scored in last chunk: {k.value}; "
          f"positives {int(d32.sum())}/{n}")
    assert err < 1e-2
    assert torch.equal(dtc, d32)
    lh, dh = tc.score_host(pcm)
    assert np.array_equal(dh, d32.cpu().numpy())


def test_stream_tensor_path(cuda_device, xiaoa_sd):
    import ww_b200

    rng = np.random.default_rng(3)
    pcm = np.clip(np.round(rng.normal(0, 0.05, 16000 * 8) * 32767), -32768, 32767).astype(np.int16)
    x = torch.from_numpy(pcm).to(cuda_device)
    f32, l32 = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn="python", cnn_impl="fp32").score(x)
    ftc, ltc = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn="python", cnn_impl="tensor").score(x)
    torch.cuda.synchronize()
    assert torch.equal(f32, ftc)
    assert (l32 - ltc).abs().max().item() < 1e-2
