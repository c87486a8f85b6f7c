"""Out-of-bounds writes, checked without a sanitizer (compute-sanitizer is closed on this pool): every output buffer the
round's new kernels write -- logits / decisions of the three clip-path hand-overs, the CTC workspace, loss and gradient of
the one-thread-per-utterance and the split wide-vocabulary kernels -- is a view into a larger allocation whose borders
carry a sentinel; the borders must be intact after the call, for sizes that leave ragged tails in every kernel."""
import numpy as np
import pytest
import torch

from oracle import mfcc as om

pytestmark = pytest.mark.gpu

GUARD = 4096  # elements on each side


def _guarded(n, dtype, device, fill):
    whole = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device=device)
    return whole, whole[GUARD:GUARD + n]


def _intact(whole, n, fill):
    lo, hi = whole[:GUARD], whole[GUARD + n:]
    if whole.dtype.is_floating_point:
        return bool((lo == fill).all() and (hi == fill).all())
    return bool((lo == fill).all() and (hi == fill).all())


@pytest.mark.parametrize("mode", ["default", "l2_chunks", "one_kernel"])
def test_clip_path_writes_stay_inside_logits_and_decisions(cuda_device, xiaoa_sd, mode):
    import ww_b200
    from ww_b200 import _lib as L

    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    sc._prep()
    ctx = sc.ctx
    opts = {"default": [(L.OPT_FUSED, 0), (L.OPT_L2_CHUNK_CLIPS, 0)], "l2_chunks": [(L.OPT_FUSED, 0), (L.OPT_L2_CHUNK_CLIPS, 4096)],
            "one_kernel": [(L.OPT_FUSED, 2), (L.OPT_L2_CHUNK_CLIPS, 0)]}[mode]
    for o, v in opts:
        ctx.check(ctx.lib.ww_set_option(ctx.h, o, v), "ww_set_option")
    try:
        for n in (1, 13, 4097, 9001):
            base = om.synth_clips_int16(min(n, 512), seed=n)
            pcm = torch.from_numpy(base).to(cuda_device).repeat((n + base.shape[0] - 1) // base.shape[0], 1)[:n].contiguous()
            lw, lg = _guarded(n, torch.float32, cuda_device, -777.0)
            dw, dc = _guarded(n, torch.uint8, cuda_device, 0xA5)
            ctx.check(ctx.lib.ww_score_clips(ctx.h, L.ptr(pcm), L.PCM_S16, n, L.CMVN_PY, L.DECIDE_LOGIT, 0.0, L.CNN_TENSOR,
                                             L.ptr(lg), L.ptr(dc), L.cur_stream(cuda_device)), "ww_score_clips")
            torch.cuda.synchronize()
            assert _intact(lw, n, -777.0) and _intact(dw, n, 0xA5), (mode, n)
            assert bool(torch.isfinite(lg).all()) and bool((lg != -777.0).all()) and bool((dc <= 1).all())
    finally:
        for o in (L.OPT_FUSED, L.OPT_L2_CHUNK_CLIPS):
            ctx.check(ctx.lib.ww_set_option(ctx.h, o, 0), "ww_set_option")


@pytest.mark.parametrize("T,B,C,S", [(63, 1000, 3, 2), (17, 129, 8, 3), (90, 7, 100, 63), (120, 9, 64, 20), (33, 5, 20, 7)])
def test_ctc_loss_writes_stay_inside_workspace_loss_and_gradient(cuda_device, T, B, C, S):
    """keyword shapes (one thread per utterance), wide vocabulary (beta + rows) and the mid-size kernels"""
    import ww_b200  # noqa: F401
    from ww_b200 import _lib as L

    rng = np.random.default_rng(T + B)
    x = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = torch.from_numpy((x - np.log(np.exp(x).sum(-1, keepdims=True))).astype(np.float32)).to(cuda_device)
    tg = torch.from_numpy(rng.integers(1, C, size=(B, S)).astype(np.int32)).to(cuda_device)
    il = torch.from_numpy(rng.integers(0, T + 1, size=B).astype(np.int32)).to(cuda_device)
    tl = torch.from_numpy(rng.integers(0, S + 1, size=B).astype(np.int32)).to(cuda_device)
    eng = L.get_context(0)
    nbytes = int(eng.lib.ww_ctc_loss_workspace_bytes(T, B, S))
    ww_, ws = _guarded(nbytes, torch.uint8, cuda_device, 0xA5)
    nw, nll = _guarded(B, torch.float32, cuda_device, -777.0)
    gw, grad = _guarded(T * B * C, torch.float32, cuda_device, -777.0)
    go = torch.ones((B,), dtype=torch.float32, device=cuda_device)
    sp = L.cur_stream(cuda_device)
    eng.check(eng.lib.ww_ctc_loss_fwd(eng.h, L.ptr(lp), lp.stride(0), lp.stride(1), T, B, C, L.ptr(tg), S, L.ptr(il),
                                      L.ptr(tl), 0, 1, L.ptr(nll), L.ptr(ws), sp), "ww_ctc_loss_fwd")
    eng.check(eng.lib.ww_ctc_loss_bwd(eng.h, L.ptr(lp), lp.stride(0), lp.stride(1), T, B, C, L.ptr(tg), S, L.ptr(il),
                                      L.ptr(tl), 0, 1, L.ptr(go), L.ptr(ws), L.ptr(grad), B * C, C, sp), "ww_ctc_loss_bwd")
    torch.cuda.synchronize()
    assert _intact(ww_, nbytes, 0xA5), "workspace overrun"
    assert _intact(nw, B, -777.0), "nll overrun"
    assert _intact(gw, T * B * C, -777.0), "gradient overrun"
    assert bool((grad != -777.0).all()), "every gradient element is written"
    assert bool(torch.isfinite(nll).all())
