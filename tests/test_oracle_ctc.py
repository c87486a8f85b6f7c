"""CPU: pin the CTC oracle against decoders lifted from the reference and against torch's CTC loss."""
import os

import numpy as np
import pytest

from oracle import ctc


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ctc_decode.npz"))


@pytest.mark.parametrize("i", [0, 1, 2])
def test_decoders_match_reference_functions(gold, i):
    lp = gold[f"lp{i}"]
    Cn = lp.shape[-1]
    chars = {k: chr(ord("a") + k - 1) if k else "_" for k in range(Cn)}
    keep = [ctc.ctc_greedy_decode(lp[b], chars) for b in range(lp.shape[0])]
    assert keep == list(gold[f"keep{i}"])
    assert ctc.decode_predictions(lp, chars) == list(gold[f"collapse{i}"])


def test_semantics_differ_on_repeats():
    lp = np.log(np.array([[.1, .8, .1], [.1, .8, .1], [.8, .1, .1], [.1, .1, .8], [.1, .1, .8]], np.float32))
    assert ctc.greedy_labels(lp, ctc.MODE_KEEP_REPEATS) == [1, 1, 2, 2]
    assert ctc.greedy_labels(lp, ctc.MODE_COLLAPSE) == [1, 2]
    assert ctc.keyword_hit([1, 1, 2, 2], [1, 2]) and not ctc.keyword_hit([1, 1], [2])
    assert ctc.detect_confidence([1, 2], [1, 2]) == 0.9


@pytest.mark.parametrize("T,B,C,S", [(63, 8, 3, 2), (20, 4, 6, 5), (30, 3, 5, 0)])
def test_numpy_ctc_matches_torch(T, B, C, S):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(T, B, C)).astype(np.float32)
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    tg = rng.integers(1, C, size=(B, max(S, 1)))
    if S >= 2:
        tg[0, 1] = tg[0, 0]  # a repeated label
    il = rng.integers(max(2 * S + 1, 1), T + 1, size=B)
    tl = rng.integers(0, S + 1, size=B)
    loss, grad = ctc.ctc_loss_torch(lp, tg, il, tl, reduction="none")
    nll, g64 = ctc.ctc_loss_numpy64(lp, tg, il, tl)
    np.testing.assert_allclose(nll, loss, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g64, grad, atol=2e-5)


def test_infinite_loss_and_zero_infinity():
    lp = np.log(np.full((2, 1, 3), 1 / 3, np.float32))
    tg = np.array([[1, 1, 2]])
    loss, _ = ctc.ctc_loss_torch(lp, tg, [2], [3], reduction="none", want_grad=False)
    assert np.isinf(loss[0])
    loss0, _ = ctc.ctc_loss_torch(lp, tg, [2], [3], reduction="none", zero_infinity=True, want_grad=False)
    assert loss0[0] == 0
    nll, _ = ctc.ctc_loss_numpy64(lp, tg, [2], [3])
    assert np.isinf(nll[0])
