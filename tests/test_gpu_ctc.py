"""GPU parity: CTC greedy decode (both reference semantics), keyword match, CTC loss forward/backward."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc as octc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", [0, 1, 2])
def test_decoders_against_reference_goldens(cuda_device, golden_dir, i):
    import ww_b200

    g = np.load(os.path.join(golden_dir, "ctc_decode.npz"))
    lp = torch.from_numpy(g[f"lp{i}"]).to(cuda_device)
    Cn = lp.shape[-1]
    chars = {k: chr(ord("a") + k - 1) if k else "_" for k in range(Cn)}
    keep = [ww_b200.ctc_greedy_decode(lp[b], chars) for b in range(lp.shape[0])]
    assert keep == list(g[f"keep{i}"])
    assert ww_b200.ctc_greedy_decode(lp[:1], chars) == g[f"keep{i}"][0]  # [1, T, C] form
    assert ww_b200.decode_predictions(lp, chars) == list(g[f"collapse{i}"])


@pytest.mark.parametrize("B,T,C", [(1, 1, 2), (37, 63, 3), (5, 100, 40), (3, 33, 4096), (2, 64, 33), (4, 257, 68), (7, 200, 130)])
@pytest.mark.parametrize("mode", ["collapse", "keep_repeats"])
def test_greedy_batch_vs_oracle(cuda_device, B, T, C, mode):
    import ww_b200

    rng = np.random.default_rng(B * 1000 + T)
    x = rng.normal(size=(B, T, C)).astype(np.float32)
    x[..., 0] += 1.0
    x = np.ascontiguousarray(np.repeat(x[:, ::2], 2, axis=1)[:, :T])  # runs of equal argmax
    lengths = rng.integers(0, T + 1, size=B).astype(np.int32)
    lengths[0] = T
    kw = [1, 2] if C > 2 else [1]
    labels, n, hits = ww_b200.greedy_batch(torch.from_numpy(x).to(cuda_device), mode=mode,
                                           lengths=torch.from_numpy(lengths), keyword=kw)
    torch.cuda.synchronize()
    labels, n, hits = labels.cpu().numpy(), n.cpu().numpy(), hits.cpu().numpy()
    m = octc.MODE_COLLAPSE if mode == "collapse" else octc.MODE_KEEP_REPEATS
    for b in range(B):
        want = octc.greedy_labels(x[b], m, length=int(lengths[b]))
        assert n[b] == len(want)
        assert labels[b, : n[b]].tolist() == want
        assert (labels[b, n[b]:] == 0).all()
        assert bool(hits[b]) == octc.keyword_hit(want, kw)
    # time-major view gives the same answer
    l2, n2, _ = ww_b200.greedy_batch(torch.from_numpy(x).to(cuda_device).transpose(0, 1).contiguous(), mode=mode,
                                     lengths=torch.from_numpy(lengths), batch_first=False)
    assert np.array_equal(l2.cpu().numpy(), labels) and np.array_equal(n2.cpu().numpy(), n)


def _greedy_raw(ctx, lp, mode, lengths, kw):
    """ww_ctc_greedy through an explicit context (the A/B switch WW_GREEDY_GENERIC is read by ww_create)."""
    from ww_b200 import _lib as L

    B, T, C = lp.shape
    labels = torch.full((B, T), -7, dtype=torch.int32, device=lp.device)
    n = torch.full((B,), -7, dtype=torch.int32, device=lp.device)
    hits = torch.full((B,), 9, dtype=torch.uint8, device=lp.device)
    kwt = torch.as_tensor(kw, dtype=torch.int32, device=lp.device)
    dm = L.DECODE_COLLAPSE if mode == "collapse" else L.DECODE_KEEP_REPEATS
    ctx.check(ctx.lib.ww_ctc_greedy(ctx.h, L.ptr(lp), lp.stride(1), lp.stride(0), T, B, C, L.ptr(lengths), dm, L.ptr(labels),
                                    L.ptr(n), L.ptr(kwt), len(kw), L.ptr(hits), L.cur_stream(lp.device)), "ww_ctc_greedy")
    torch.cuda.synchronize()
    return labels.cpu().numpy(), n.cpu().numpy(), hits.cpu().numpy()


@pytest.mark.parametrize("B,T,C,kw", [(1000, 63, 1, [1, 1, 1]), (513, 63, 3, [1, 2]), (77, 64, 4, [3]), (129, 40, 2, [1] * 9),
                                      (64, 7, 3, [2, 1]), (5, 1, 1, [1]), (4099, 33, 3, [])])
@pytest.mark.parametrize("mode", ["collapse", "keep_repeats"])
def test_greedy_keyword_shapes(cuda_device, B, T, C, kw, mode):
    """T <= 64, C <= 4 run ctc_greedy_short_kernel: same labels, lengths and keyword hits as the generic kernel
    (WW_GREEDY_GENERIC=1 context) and as the oracle; C = 1 is the binary posterior in logit form (label 1 iff x > 0)."""
    from ww_b200 import _lib as L

    rng = np.random.default_rng(B * 131 + T * 7 + C)
    x = rng.normal(size=(B, T, C)).astype(np.float32)
    if C > 1:
        x[..., 0] += 0.5
    x = np.ascontiguousarray(np.repeat(x[:, ::2], 2, axis=1)[:, :T])  # runs of equal argmax
    x[B // 2] = 0.0                                                   # ties / exact zeros
    lengths = rng.integers(0, T + 1, size=B).astype(np.int32)
    lengths[0] = T
    lp = torch.from_numpy(x).to(cuda_device)
    ln = torch.from_numpy(lengths).to(cuda_device)
    os.environ["WW_GREEDY_GENERIC"] = "1"
    try:
        generic = L.Context(cuda_device.index or 0)
    finally:
        del os.environ["WW_GREEDY_GENERIC"]
    try:
        for lens in (ln, None):
            got = _greedy_raw(L.get_context(cuda_device.index or 0), lp, mode, lens, kw)
            ref = _greedy_raw(generic, lp, mode, lens, kw)
            for g, r in zip(got, ref):
                assert np.array_equal(g, r)
            labels, n, hits = got
            m = octc.MODE_COLLAPSE if mode == "collapse" else octc.MODE_KEEP_REPEATS
            x2 = x if C > 1 else np.concatenate([np.zeros_like(x), x], axis=-1)
            for b in range(0, B, max(1, B // 97)):
                want = octc.greedy_labels(x2[b], m, length=int(lengths[b]) if lens is not None else T)
                assert n[b] == len(want) and labels[b, : n[b]].tolist() == want and (labels[b, n[b]:] == 0).all()
                assert bool(hits[b]) == octc.keyword_hit(want, kw)
    finally:
        generic.close()


def test_argmax_ties_take_first_index(cuda_device):
    import ww_b200

    lp = torch.zeros(2, 8, 40, device=cuda_device)
    lp[0, :, 5] = 1.0
    lp[0, :, 9] = 1.0
    labels, n, _ = ww_b200.greedy_batch(lp, mode="keep_repeats")
    assert labels[0, :8].tolist() == [5] * 8 and int(n[1]) == 0
    lp3 = torch.zeros(1, 4, 3, device=cuda_device)
    lp3[0, :, 1:] = 2.0
    labels, n, _ = ww_b200.greedy_batch(lp3, mode="collapse")
    assert labels[0, : int(n[0])].tolist() == [1]
    want = torch.max(lp3.cpu()[0], dim=1)[1].tolist()
    assert want == [1, 1, 1, 1]


def test_keyword_detector(cuda_device):
    import ww_b200

    c2i = {"_": 0, "x": 1, "a": 2}
    det = ww_b200.CTCKeywordDetector(c2i, ["xa", "ax"], threshold=0.8)
    seq = [0, 1, 1, 0, 2, 0, 0]
    lp = torch.full((2, 7, 3), -5.0, device=cuda_device)
    for t, s in enumerate(seq):
        lp[0, t, s] = 0.0
    lp[1, :, 0] = 0.0
    assert det.ctc_greedy_decode(lp[0]) == "xxa"
    res = det.detect_batch(lp)
    assert res[0] == [("xa", 0.9)] and res[1] == []
    assert det.calculate_confidence("xxa", "xa") == 0.9 and det.calculate_confidence("xx", "xa") == 0.0


def test_keyword_detector_reference_signature_and_streaming_loop(cuda_device):
    """CTCKeywordDetector(model, char_to_idx, keywords) and detect_keywords(stream) as in ml_models/test.py:158-200:
    the oracle is the reference loop restated on the CPU with the same stand-in model and the same features."""
    import ww_b200
    from oracle import ctc as octc

    c2i = {"_": 0, "x": 1, "a": 2}
    torch.manual_seed(0)
    lin = torch.nn.Linear(13, 3).to(cuda_device)

    class Model(torch.nn.Module):
        def forward(self, feats):                      # [1, T, 13] -> log-probs [1, T, 3]
            return torch.log_softmax(lin(feats) * 0.5, dim=-1)

    det = ww_b200.CTCKeywordDetector(Model(), c2i, ["xa", "ax", "a"], threshold=0.8)
    assert det.model is not None and det.keywords == ["xa", "ax", "a"]
    g = torch.Generator().manual_seed(1)
    stream = [torch.randn(1, 4000, generator=g) * 0.1 for _ in range(12)]
    got = det.detect_keywords(stream)
    # the reference loop, chunk by chunk (features of buffer[0], slide by five)
    want, buffer = [], []
    i2c = {v: k for k, v in c2i.items()}
    for chunk in stream:
        buffer.append(chunk)
        feats = ww_b200.mfcc_batch(buffer[0][0][None].to(cuda_device))[0].T
        lp = torch.log_softmax(lin(feats) * 0.5, dim=-1).detach().cpu().numpy()
        text = "".join(i2c[i] for i in octc.greedy_labels(lp, octc.MODE_KEEP_REPEATS))
        want += [(kw, 0.9) for kw in ["xa", "ax", "a"] if kw in text]
        buffer = buffer[5:]
    assert got == want and len(got) > 0
    with pytest.raises(ValueError):
        ww_b200.CTCKeywordDetector(c2i, ["xa"]).detect_keywords(stream)
    with pytest.raises(TypeError):
        ww_b200.CTCKeywordDetector(c2i)


def _rand_problem(T, B, C, S, seed, full_len=False):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    tg = rng.integers(1, C, size=(B, max(S, 1))).astype(np.int64)
    if S >= 2:
        tg[0, 1] = tg[0, 0]
    il = np.full(B, T) if full_len else rng.integers(min(2 * S + 1, T), T + 1, size=B)
    tl = np.full(B, S) if full_len else rng.integers(0, S + 1, size=B)
    return lp.astype(np.float32), tg, il, tl


@pytest.mark.parametrize("T,B,C,S", [(63, 64, 3, 1), (63, 200, 3, 2), (50, 9, 20, 7), (120, 5, 50, 40), (801, 2, 300, 32),
                                     (200, 4, 100, 63), (300, 3, 200, 70), (37, 6, 64, 15), (5, 3, 8, 4)])
@pytest.mark.parametrize("reduction", ["mean", "none"])
def test_ctc_loss_fwd_bwd_vs_torch(cuda_device, T, B, C, S, reduction):
    """loss rtol 1e-5, grad atol 1e-5 against torch.nn.functional.ctc_loss on CPU.

    With reduction='none' and long inputs the per-sample nll reaches 1e3..1e4, where fp32 alpha/beta (ulp
    ~2e-4..5e-4) limit BOTH fp32 implementations to ~5e-4 relative on the target-class gradients; there the
    arbiter is torch's fp64 result and the kernel must be as close to it as torch's own fp32 path is."""
    import ww_b200

    lp, tg, il, tl = _rand_problem(T, B, C, S, seed=T + B)
    want_loss, want_grad = octc.ctc_loss_torch(lp, tg, il, tl, reduction=reduction)
    x = torch.from_numpy(lp).to(cuda_device).requires_grad_(True)
    crit = ww_b200.CTCLoss(blank=0, reduction=reduction)
    loss = crit(x, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    (loss.sum() if reduction == "none" else loss).backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want_loss, rtol=1e-5, atol=1e-5)
    got = x.grad.cpu().numpy()
    if reduction == "mean":
        np.testing.assert_allclose(got, want_grad, atol=1e-5)
    else:
        _, g64 = octc.ctc_loss_torch(lp, tg, il, tl, reduction=reduction, dtype="float64")
        err_mine = np.abs(got - g64).max()
        err_torch32 = np.abs(want_grad - g64).max()
        print(f"grad err vs fp64: kernel {err_mine:.2e}, torch fp32 {err_torch32:.2e}")
        assert err_mine <= max(1e-5, 3.0 * err_torch32)


@pytest.mark.parametrize("T,B,C,S", [(63, 1000, 3, 3), (40, 333, 8, 3), (63, 257, 5, 2), (17, 129, 2, 1), (63, 4097, 3, 2),
                                     (801, 33, 5, 3)])
@pytest.mark.parametrize("blank", [0, 1])
def test_ctc_loss_keyword_shapes_one_thread_per_utterance(cuda_device, T, B, C, S, blank):
    """The S <= 3, C <= 8 kernels (one thread per utterance, time-major alpha): ragged input lengths down to 0, target
    lengths 0..S, repeated labels, a non-zero blank index, batches that do not fill the last CTA; loss and gradient
    against torch.nn.functional.ctc_loss (reduction 'none': every utterance's loss and its own gradient rows)."""
    import ww_b200

    rng = np.random.default_rng(T * 1000 + B + C + S + blank)
    x = rng.normal(size=(T, B, C)).astype(np.float32) * 2.0
    lp = (x - np.log(np.exp(x).sum(-1, keepdims=True))).astype(np.float32)
    labels = np.array([c for c in range(C) if c != blank])
    tg = labels[rng.integers(0, len(labels), size=(B, S))].astype(np.int64)
    tl = rng.integers(0, S + 1, size=B)
    il = rng.integers(0, T + 1, size=B)
    il[: B // 2] = T
    if S >= 2:
        tg[::3, 1] = tg[::3, 0]                      # repeated labels need the blank between them
    want_loss, want_grad = octc.ctc_loss_torch(lp, tg, il, tl, blank=blank, reduction="none", zero_infinity=True)
    xg = torch.from_numpy(lp).to(cuda_device).requires_grad_(True)
    loss = ww_b200.ctc_loss(xg, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl), blank=blank,
                            reduction="none", zero_infinity=True)
    w = torch.from_numpy(rng.uniform(0.5, 2.0, size=B).astype(np.float32)).to(cuda_device)
    (loss * w).sum().backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want_loss, rtol=1e-5, atol=1e-5)
    # per-utterance losses reach ~1e2 here, where fp32 alpha / beta limit both fp32 implementations: the arbiter is
    # torch's fp64 gradient, and the kernel must be as close to it as torch's own fp32 path is
    _, g64 = octc.ctc_loss_torch(lp, tg, il, tl, blank=blank, reduction="none", zero_infinity=True, dtype="float64")
    wn = w.cpu().numpy()[None, :, None]
    err_mine = np.abs(xg.grad.cpu().numpy() - g64 * wn).max()
    err_torch32 = np.abs(want_grad * wn - g64 * wn).max()
    print(f"grad err vs fp64: kernel {err_mine:.2e}, torch fp32 {err_torch32:.2e}")
    assert err_mine <= max(2e-5, 3.0 * err_torch32)


@pytest.fixture
def ctc_split_mode(request, cuda_device):
    """WW_OPT_CTC_SPLIT for the duration of a test (1 is the default: beta recursion, then fill + patches per row)"""
    from ww_b200 import _lib as L

    ctx = L.get_context(0)
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_CTC_SPLIT, request.param), "ww_set_option")
    yield request.param
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_CTC_SPLIT, 1), "ww_set_option")


@pytest.mark.parametrize("ctc_split_mode", [1, 2, 3, 0], indirect=True)
@pytest.mark.parametrize("T,B,C,S", [(120, 9, 64, 20), (200, 5, 256, 40), (90, 7, 100, 63), (150, 4, 1000, 31)])
def test_ctc_loss_wide_vocabulary_split_backward(cuda_device, T, B, C, S, ctc_split_mode):
    """C >= 64, 2S+1 <= 128: beta recursion + row-parallel gradient.  Targets with many repeated labels (adjacent and
    not), a label equal to the blank index... is not legal for torch, so the blank-index fold is covered by repeated
    labels only; ragged lengths, an infeasible utterance, zero-length targets; a second backward through the same graph
    (the workspace's alpha must survive the first)."""
    import ww_b200

    rng = np.random.default_rng(T + B + C + S)
    x = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = (x - np.log(np.exp(x).sum(-1, keepdims=True))).astype(np.float32)
    tg = rng.integers(1, min(C, 6), size=(B, S)).astype(np.int64)       # five labels only: repeats everywhere
    tg[1] = rng.integers(1, C, size=S)                                  # one utterance without repeats (fast path)
    tl = rng.integers(1, S + 1, size=B)
    tl[0] = S
    tl[-1] = 0
    il = np.full(B, T)
    il[2] = T // 2
    il[3] = 3                                                           # infeasible for its target
    tl[3] = S
    want_loss, want_grad = octc.ctc_loss_torch(lp, tg, il, tl, reduction="none", zero_infinity=True)
    _, g64 = octc.ctc_loss_torch(lp, tg, il, tl, reduction="none", zero_infinity=True, dtype="float64")
    xg = torch.from_numpy(lp).to(cuda_device).requires_grad_(True)
    loss = ww_b200.ctc_loss(xg, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl), reduction="none",
                            zero_infinity=True)
    loss.sum().backward(retain_graph=True)
    g1 = xg.grad.clone()
    xg.grad = None
    loss.sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(g1, xg.grad)                                     # deterministic, and alpha was not consumed
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want_loss, rtol=1e-5, atol=1e-5)
    err_mine = np.abs(g1.cpu().numpy() - g64).max()
    err_torch32 = np.abs(want_grad - g64).max()
    print(f"grad err vs fp64: kernel {err_mine:.2e}, torch fp32 {err_torch32:.2e}")
    assert err_mine <= max(1e-5, 3.0 * err_torch32)


@pytest.mark.parametrize("T,B,C,S", [(300, 9, 64, 20), (260, 5, 256, 40), (257, 7, 100, 63), (256, 33, 4096, 32), (120, 9, 64, 20)])
def test_ctc_beta_recursion_beside_the_forward_pass_gives_the_same_bits(cuda_device, T, B, C, S):
    """WW_CTC_BETA_IN_FWD: the beta recursion of the wide-vocabulary path runs inside the forward call, beside alpha
    (it stores beta; the rows pass forms alpha + beta and re-derives nll / the live rows) -- loss and gradient must
    equal the two-call form bit for bit, with ragged lengths, an infeasible utterance, an empty target, repeated
    labels, a second backward through the kept graph, and with and without zero_infinity."""
    import ww_b200
    from ww_b200 import ctc as wctc

    rng = np.random.default_rng(7 * T + B + C + S)
    x = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = torch.log_softmax(torch.from_numpy(x), dim=-1).to(cuda_device)
    tg = rng.integers(1, min(C, 6), size=(B, S)).astype(np.int64)
    tg[1] = rng.integers(1, C, size=S)
    tl = rng.integers(1, S + 1, size=B)
    tl[0] = S
    tl[-1] = 0
    il = np.full(B, T)
    il[2] = T // 2
    il[3] = 3
    tl[3] = S
    res = {}
    for zi in (True, False):
        for beta_in_fwd in (False, True):
            wctc.BETA_IN_FWD = beta_in_fwd
            try:
                xg = lp.clone().requires_grad_(True)
                loss = ww_b200.ctc_loss(xg, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl),
                                        reduction="none", zero_infinity=zi)
                w = torch.linspace(0.5, 1.5, B, device=cuda_device)
                (loss * w)[torch.isfinite(loss)].sum().backward(retain_graph=True)
                g1 = xg.grad.clone()
                xg.grad = None
                (loss * w)[torch.isfinite(loss)].sum().backward()
                torch.cuda.synchronize()
                # (without zero_infinity the infeasible utterance's rows are NaN, as torch's are)
                assert torch.equal(torch.nan_to_num(g1, nan=123.0), torch.nan_to_num(xg.grad, nan=123.0))
                res[(zi, beta_in_fwd)] = (loss.detach().clone(), g1)
            finally:
                wctc.BETA_IN_FWD = True
        l0, g0 = res[(zi, False)]
        l1, g1 = res[(zi, True)]
        assert torch.equal(l0, l1)
        assert torch.equal(torch.nan_to_num(g0, nan=123.0), torch.nan_to_num(g1, nan=123.0)), (zi, (g0 - g1).abs().max())


def test_ctc_loss_zero_infinity_and_blank_index(cuda_device):
    import ww_b200

    # impossible alignment: T=2 frames cannot emit "112"
    lp = torch.log(torch.full((2, 2, 3), 1 / 3, device=cuda_device)).requires_grad_(True)
    tg = torch.tensor([[1, 1, 2], [1, 0, 0]])
    il, tl = torch.tensor([2, 2]), torch.tensor([3, 1])
    inf = ww_b200.ctc_loss(lp, tg, il, tl, reduction="none", zero_infinity=False)
    assert torch.isinf(inf[0]) and torch.isfinite(inf[1])
    z = ww_b200.ctc_loss(lp, tg, il, tl, reduction="none", zero_infinity=True)
    z.sum().backward()
    want, wg = octc.ctc_loss_torch(lp.detach().cpu().numpy(), tg.numpy(), il.numpy(), tl.numpy(), reduction="none",
                                   zero_infinity=True)
    np.testing.assert_allclose(z.detach().cpu().numpy(), want, atol=1e-6)
    np.testing.assert_allclose(lp.grad.cpu().numpy(), wg, atol=1e-6)
    # test.py uses a non-zero blank index (char_to_idx['_'])
    lp2, tg2, il2, tl2 = _rand_problem(30, 6, 4, 3, seed=1)
    tg2 = np.where(tg2 == 2, 3, tg2)  # blank = 2 must not appear in targets
    want, wg = octc.ctc_loss_torch(lp2, tg2, il2, tl2, blank=2, reduction="mean")
    x = torch.from_numpy(lp2).to(cuda_device).requires_grad_(True)
    loss = ww_b200.CTCLoss(blank=2, zero_infinity=True)(x, torch.from_numpy(tg2), torch.from_numpy(il2),
                                                       torch.from_numpy(tl2))
    loss.backward()
    np.testing.assert_allclose(loss.item(), want, rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), wg, atol=1e-5)
