"""GPU parity: the tcgen05 path returns the fp32 path's decisions for ANY loaded weights and ANY input.

The guard band is calibrated by ww_load_weights for the weights just loaded (ww_tc_band_info); inside it the exact fp32
kernel decides.  Decision rules of the reference: sigmoid(out) > 0.5 (ml_models/main.py:52-53) and
sigmoid(out) * 100 >= 80 (esp_wake_word_detector.cpp:226-228,245).
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LN4 = math.log(4.0)


def _sd(rng, num_classes=1, scale=1.0):
    return {
        "conv_layers.0.weight": (rng.normal(0, 0.2, (32, 13, 3)) * scale).astype(np.float32),
        "conv_layers.3.weight": (rng.normal(0, 0.1, (64, 32, 3)) * scale).astype(np.float32),
        "conv_layers.6.weight": (rng.normal(0, 0.1, (128, 64, 3)) * scale).astype(np.float32),
        "classifier.0.weight": (rng.normal(0, 0.1, (64, 128)) * scale).astype(np.float32),
        "classifier.2.weight": (rng.normal(0, 0.2, (num_classes, 64)) * scale).astype(np.float32),
    }


def _forward(sd, x, impl, cmvn, decide, thr):
    """ww_cnn_forward over [n,13,63] windows -> (logits, decisions, re-scored count)."""
    import ww_b200
    from ww_b200 import _lib as L
    from ww_b200.model import _push_weights, _tokens

    ctx = L.get_context(0)
    _push_weights(ctx, sd, ("band-test", next(_tokens)))
    n = x.shape[0]
    logits = torch.empty((n, ctx.num_classes), device=x.device)
    dec = torch.zeros((n,), dtype=torch.uint8, device=x.device)
    ctx.check(ctx.lib.ww_cnn_forward(ctx.h, L.ptr(x), 819, 63, 1, n, cmvn, decide, thr,
                                     L.CNN_TENSOR if impl == "tensor" else L.CNN_FP32, L.ptr(logits),
                                     L.ptr(dec) if decide != L.DECIDE_NONE else None, L.cur_stream(x.device)), "fwd")
    torch.cuda.synchronize()
    k = C.c_int(0)
    if impl == "tensor":
        ctx.check(ctx.lib.ww_debug_tc(ctx.h, None, C.byref(k)), "dbg")
    return logits, dec, k.value, ctx.tc_band_info()


def _adversarial_windows(rng):
    """impulse / full-scale square / single hot frame / CMVN extreme point / silence-like rows, [n,13,63] fp32."""
    w = []
    for t in (0, 1, 31, 61, 62):                       # one hot frame: after CMVN |z| = 62/sqrt(63) = 7.81
        x = np.zeros((13, 63), np.float32)
        x[:, t] = rng.normal(0, 5, 13)
        w.append(x)
    for per in (1, 2, 3, 7, 16, 31):                   # full-scale square waves
        x = np.where((np.arange(63) // per) % 2 == 0, 7.8, -7.8).astype(np.float32)
        w.append(np.tile(x, (13, 1)) * rng.choice([-1.0, 1.0], (13, 1)).astype(np.float32))
    for q in range(13):                                # a single impulse in one coefficient
        x = np.zeros((13, 63), np.float32)
        x[q, rng.integers(63)] = 1.0
        w.append(x)
    x = np.full((13, 63), -0.125988, np.float32)       # the CMVN extreme point itself
    x[np.arange(13), rng.integers(0, 63, 13)] = 7.811249
    w.append(x)
    x = np.zeros((13, 63), np.float32)                 # silence: c0 = -87.377, the rest 0
    x[0] = -87.377
    w.append(x)
    w.append(np.zeros((13, 63), np.float32))
    w.append(rng.normal(0, 1e-4, (13, 63)).astype(np.float32))    # tiny values (fp16 subnormals)
    w.append(rng.normal(0, 300.0, (13, 63)).astype(np.float32))   # huge values
    return np.stack(w)


def _bisect_to_threshold(sd, a, b, thr, dev, iters=40):
    """Windows on the segment a..b whose fp32 logit is as close to `thr` as fp32 allows (vectorised bisection with the
    exact kernel).  a: logits < thr, b: logits > thr."""
    from ww_b200 import _lib as L

    lo = torch.zeros(a.shape[0], device=dev)
    hi = torch.ones(a.shape[0], device=dev)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        x = (a + mid[:, None, None] * (b - a)).contiguous()
        lg = _forward(sd, x, "fp32", L.CMVN_NONE, L.DECIDE_NONE, 0.0)[0][:, 0]
        up = lg > thr
        hi = torch.where(up, mid, hi)
        lo = torch.where(up, lo, mid)
    return lo, hi


@pytest.mark.parametrize("which", ["xiaoa", "random", "random_c3", "random_x3", "random_x10"])
def test_decisions_equal_fp32_on_windows_at_the_threshold(cuda_device, xiaoa_sd, which):
    """Thousands of windows whose exact logit sits within a few fp32 ulps of a threshold, on both sides, plus the
    adversarial shapes: tensor-path decisions and thresholded logits must be the fp32 path's, for the shipped weights
    and for random / 3x / 10x-scaled ones (10x per layer = 1e5 on the logits: the band must scale or the path switch off)."""
    from ww_b200 import _lib as L

    rng = np.random.default_rng(len(which))
    sd = xiaoa_sd if which == "xiaoa" else _sd(rng, 3 if which == "random_c3" else 1,
                                              {"random_x3": 3.0, "random_x10": 10.0}.get(which, 1.0))
    n = 4096
    x = torch.from_numpy(rng.normal(0, 1, (n, 13, 63)).astype(np.float32)).to(cuda_device)
    l32 = _forward(sd, x, "fp32", L.CMVN_NONE, L.DECIDE_NONE, 0.0)[0][:, 0]
    for thr in (0.0, LN4):
        below, above = x[l32 < thr], x[l32 > thr]
        m = min(below.shape[0], above.shape[0], 1500)
        if m == 0:
            # every random window lies on one side: mirror the set through the origin is not possible for a ReLU
            # network, so scale-shift instead -- the adversarial set below still exercises the band
            continue
        a, b = below[:m], above[:m]
        lo, hi = _bisect_to_threshold(sd, a, b, thr, cuda_device)
        edge = torch.cat([a + lo[:, None, None] * (b - a), a + hi[:, None, None] * (b - a)]).contiguous()
        for decide, t_arg in ((L.DECIDE_LOGIT, thr), (L.DECIDE_DEVICE, 100.0 / (1.0 + math.exp(-thr)))):
            g32, d32, _, _ = _forward(sd, edge, "fp32", L.CMVN_NONE, decide, t_arg)
            gtc, dtc, k, info = _forward(sd, edge, "tensor", L.CMVN_NONE, decide, t_arg)
            assert torch.equal(dtc, d32), f"{which} thr={thr} decide={decide}: {(dtc != d32).sum().item()} flips"
            if info["enabled"]:
                assert k >= 0.9 * edge.shape[0]        # windows AT the threshold are inside the band: re-scored
        # DECIDE_NONE (LightweightKWS.forward): thresholding the returned logits reproduces the fp32 decisions
        g32 = _forward(sd, edge, "fp32", L.CMVN_NONE, L.DECIDE_NONE, 0.0)[0]
        gtc, _, k, info = _forward(sd, edge, "tensor", L.CMVN_NONE, L.DECIDE_NONE, 0.0)
        assert torch.equal(gtc[:, 0] > thr, g32[:, 0] > thr) and torch.equal(gtc[:, 0] >= thr, g32[:, 0] >= thr)
        print(f"{which} thr={thr:.3f}: {edge.shape[0]} windows at the threshold, {k} re-scored, band info {info}")
    # away from the thresholds the tensor logits stay inside the calibrated bound beta * ||x||
    gtc, _, _, info = _forward(sd, x, "tensor", L.CMVN_NONE, L.DECIDE_NONE, 0.0)
    g32 = _forward(sd, x, "fp32", L.CMVN_NONE, L.DECIDE_NONE, 0.0)[0]
    ratio = ((gtc - g32).abs().max(dim=1)[0] / x.flatten(1).norm(dim=1)).max().item()
    print(f"{which}: max |tc - fp32| / ||x|| = {ratio:.3e} against beta = {info['beta']:.3e} "
          f"(calibrated {info['beta_calibrated']:.3e}, rigorous {info['beta_rigorous']:.3e}, enabled {info['enabled']})")
    assert ratio <= info["beta"] or not info["enabled"]
    assert info["beta"] <= info["beta_rigorous"]


@pytest.mark.parametrize("cmvn", ["none", "python", "device"])
def test_adversarial_inputs(cuda_device, xiaoa_sd, cmvn):
    from ww_b200 import _lib as L

    rng = np.random.default_rng(11)
    x = torch.from_numpy(_adversarial_windows(rng)).to(cuda_device)
    mode = {"none": L.CMVN_NONE, "python": L.CMVN_PY, "device": L.CMVN_DEVICE}[cmvn]
    for decide, thr in ((L.DECIDE_LOGIT, 0.0), (L.DECIDE_DEVICE, 80.0), (L.DECIDE_NONE, 0.0)):
        g32, d32, _, _ = _forward(xiaoa_sd, x, "fp32", mode, decide, thr)
        gtc, dtc, k, info = _forward(xiaoa_sd, x, "tensor", mode, decide, thr)
        assert torch.isfinite(gtc).all()
        assert torch.equal(dtc, d32)
        assert torch.equal(gtc[:, 0] > 0, g32[:, 0] > 0) and torch.equal(gtc[:, 0] >= LN4, g32[:, 0] >= LN4)
        nx = x.flatten(1).norm(dim=1) if cmvn == "none" else torch.full((x.shape[0],), 43.0, device=cuda_device)
        err = (gtc - g32).abs().max(dim=1)[0]
        assert (err <= info["beta"] * nx + 1e-6).all(), (cmvn, decide, err.max().item())
    print(f"cmvn={cmvn}: {x.shape[0]} adversarial windows, {k} re-scored in the last launch")


def test_overflowing_weights_switch_the_tensor_path_off(cuda_device):
    """Weights whose activations leave the fp16 range: calibration reports it and `tensor` silently IS the fp32 kernel."""
    from ww_b200 import _lib as L

    rng = np.random.default_rng(5)
    sd = _sd(rng, 1, scale=60.0)
    x = torch.from_numpy(rng.normal(0, 1, (512, 13, 63)).astype(np.float32)).to(cuda_device)
    g32, d32, _, _ = _forward(sd, x, "fp32", L.CMVN_PY, L.DECIDE_LOGIT, 0.0)
    gtc, dtc, _, info = _forward(sd, x, "tensor", L.CMVN_PY, L.DECIDE_LOGIT, 0.0)
    print(info)
    assert torch.equal(dtc, d32)
    assert (not info["enabled"]) or info["norm_limit"] < 28.4 or torch.equal(gtc, g32) or \
        ((gtc - g32).abs().max().item() <= info["band_python_cmvn"])


def test_multiclass_argmax_is_the_fp32_argmax(cuda_device):
    """num_classes = 3 (the CTC head of ml_models/test.py): windows whose two best classes tie within the band are
    re-scored, so the per-frame argmax of the greedy decoders does not depend on the operand precision."""
    from ww_b200 import _lib as L

    rng = np.random.default_rng(9)
    sd = _sd(rng, 3)
    x = torch.from_numpy(rng.normal(0, 1, (20000, 13, 63)).astype(np.float32)).to(cuda_device)
    g32 = _forward(sd, x, "fp32", L.CMVN_PY, L.DECIDE_NONE, 0.0)[0]
    gtc, _, k, info = _forward(sd, x, "tensor", L.CMVN_PY, L.DECIDE_NONE, 0.0)
    assert torch.equal(gtc.argmax(dim=1), g32.argmax(dim=1))
    print(f"3 classes: {k} of 20000 windows re-scored, band {info['band_python_cmvn']:.4f}")


def test_class_count_change_rebuilds_the_host_path(cuda_device, xiaoa_sd):
    """score_host with a 1-class model, then with a 3-class model on the same context: the pinned and
    device logit buffers are re-sized, results equal the device-buffer path."""
    import ww_b200
    from oracle import mfcc as om

    rng = np.random.default_rng(2)
    pcm = om.synth_clips_int16(700, seed=5)
    s1 = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    l1, d1 = s1.score_host(pcm)
    s3 = ww_b200.WakeWordScorer(_sd(rng, 3), device=0)
    l3, d3 = s3.score_host(pcm)
    assert l1.shape == (700, 1) and l3.shape == (700, 3)
    x = torch.from_numpy(pcm).to(cuda_device)
    l3d, d3d = s3.score(x)
    np.testing.assert_array_equal(l3, l3d.cpu().numpy())
    np.testing.assert_array_equal(d3, d3d.cpu().numpy())
    l1b, d1b = s1.score_host(pcm)                       # and back
    np.testing.assert_array_equal(l1, l1b)
    np.testing.assert_array_equal(d1, d1b)


def test_session_refuses_a_model_with_another_class_count(cuda_device, xiaoa_sd):
    from ww_b200 import _lib as L

    import ww_b200

    sess = ww_b200.StreamSession(xiaoa_sd, 2, max_chunk_samples=800, device=0)
    sess.write(np.zeros((2, 800), np.int16))
    ctx = L.get_context(0)
    from ww_b200.model import _push_weights, _tokens
    _push_weights(ctx, _sd(np.random.default_rng(0), 3), ("other", next(_tokens)))
    # C ABI level (the Python wrapper has its own owner check): the write must fail, not overflow the logit buffers
    chunk = np.zeros((2, 800), np.int16)
    rc = ctx.lib.ww_session_write(sess.h, chunk.ctypes.data_as(C.c_void_p), 800)
    assert rc == -1 and b"class count" in ctx.lib.ww_last_error(ctx.h)
    sess.close()


def test_failed_requantisation_leaves_no_half_updated_state(cuda_device, xiaoa_sd):
    """ww_quantize_weights_i8 validates the exponents before touching the previous quantisation."""
    import ww_b200
    from ww_b200 import _lib as L
    from ww_b200.model import XIAOA_EXPONENTS, _push_weights, _tokens

    ctx = L.get_context(0)
    _push_weights(ctx, xiaoa_sd, ("q", next(_tokens)), XIAOA_EXPONENTS)
    x = torch.randint(-128, 128, (64, 13, 63), dtype=torch.int8, device=cuda_device)
    out = torch.empty((64, 1), dtype=torch.int8, device=cuda_device)
    ctx.check(ctx.lib.ww_cnn_forward_i8(ctx.h, L.ptr(x), 64, L.ptr(out), L.cur_stream(cuda_device)), "i8")
    torch.cuda.synchronize()
    want = out.clone()
    bad = list(XIAOA_EXPONENTS)
    bad[2] = -40                                        # activation exponent below input + weight: negative shift
    rc = ctx.lib.ww_quantize_weights_i8(ctx.h, (C.c_int * 12)(*bad))
    assert rc == -4
    out.zero_()
    ctx.check(ctx.lib.ww_cnn_forward_i8(ctx.h, L.ptr(x), 64, L.ptr(out), L.cur_stream(cuda_device)), "i8 after failure")
    torch.cuda.synchronize()
    assert torch.equal(out, want)                       # the previous, valid quantisation is still in force


def test_concurrent_fused_calls_on_one_context_are_refused_not_raced(cuda_device, xiaoa_sd):
    """Two host threads inside ww_score_clips_host on ONE context: each call either completes with the right answer
    or returns WW_ERR_BUSY (ctypes releases the GIL for the duration of the call)."""
    import threading

    import ww_b200
    from oracle import mfcc as om
    from ww_b200 import _lib as L

    pcm = np.tile(om.synth_clips_int16(256, seed=3), (128, 1))      # 32768 clips = 1 GB
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    want_l, want_d = sc.score_host(pcm)
    results = []

    def work():
        lg = np.empty((pcm.shape[0], 1), np.float32)
        dc = np.empty((pcm.shape[0],), np.uint8)
        rc = sc.ctx.lib.ww_score_clips_host(sc.ctx.h, C.c_void_p(pcm.ctypes.data), L.PCM_S16, pcm.shape[0], sc.cmvn,
                                            sc.decide, sc.threshold, sc.cnn_impl, lg.ctypes.data_as(C.c_void_p),
                                            dc.ctypes.data_as(C.c_void_p))
        results.append((rc, lg, dc))

    ts = [threading.Thread(target=work) for _ in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    codes = sorted(r[0] for r in results)
    print("return codes", codes)
    assert set(codes) <= {0, L.ERR_BUSY} and codes.count(0) >= 1
    for rc, lg, dc in results:
        if rc == 0:
            np.testing.assert_array_equal(lg, want_l)
            np.testing.assert_array_equal(dc, want_d)
