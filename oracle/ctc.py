"""CTC oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restated reference code:
  ml_models/test.py:201-217   CTCKeywordDetector.ctc_greedy_decode
        argmax per frame; keep idx iff idx != prev_char and idx != 0, with
        prev_char FROZEN at -1 (the update at :215 is commented out) => only
        index-0 frames are dropped, repeats are NOT collapsed.
  ml_models/test.py:168-200,230-235  detect_keywords / calculate_confidence
        `keyword in decoded_text` => confidence constant 0.9 > threshold 0.8.
  ml_models/ctc.py:453-471    THCHS30Trainer.decode_predictions
        textbook best path: keep iff tok != 0 and tok != prev; prev starts 0
        and is always updated.
  ml_models/test.py:89,111-112  nn.CTCLoss(blank=..., zero_infinity=True)
  ml_models/ctc.py:369,393-401  nn.CTCLoss(blank=0)  (reduction='mean')

The three decoders cannot be imported from the reference (both scripts train
at import time and need librosa/matplotlib), so their behaviour is restated in our own
code.  The loss is pinned by PyTorch itself: `ctc_loss_torch` calls
torch.nn.functional.ctc_loss on CPU; `ctc_loss_numpy64` is an independent
alpha/beta implementation checked against it (tests/test_oracle_ctc.py).
The reference holds no golden CTC vectors.
"""
from __future__ import annotations

import numpy as np

MODE_KEEP_REPEATS = 0   # ml_models/test.py:201-217
MODE_COLLAPSE = 1       # ml_models/ctc.py:453-471


def argmax_first(log_probs):
    """argmax over the last axis, first index on ties (torch.max on CPU)."""
    return np.argmax(np.asarray(log_probs), axis=-1)


def greedy_labels(log_probs, mode, length=None):
    """log_probs: [T, C] -> list of kept label indices."""
    idx = argmax_first(log_probs)
    if length is not None:
        idx = idx[:length]
    out = []
    if mode == MODE_KEEP_REPEATS:
        prev = -1
        for i in idx.tolist():
            if i != prev and i != 0:
                out.append(i)
            # prev = i   (commented out in the reference, test.py:215)
    else:
        prev = 0
        for i in idx.tolist():
            if i != 0 and i != prev:
                out.append(i)
            prev = i
    return out


def ctc_greedy_decode(log_probs, char_list):
    """ml_models/test.py:201-217 (accepts [T,C] or [1,T,C]); returns str."""
    lp = np.asarray(log_probs)
    if lp.ndim == 3:
        lp = lp[0]
    return "".join(char_list[i] for i in greedy_labels(lp, MODE_KEEP_REPEATS))


def decode_predictions(log_probs, idx_to_char):
    """ml_models/ctc.py:453-471; log_probs [B,T,V] -> list[str]."""
    out = []
    for b in range(np.asarray(log_probs).shape[0]):
        lab = greedy_labels(np.asarray(log_probs)[b], MODE_COLLAPSE)
        out.append("".join(idx_to_char.get(i, "<unk>") for i in lab))
    return out


def keyword_hit(labels, keyword_labels):
    """`keyword in decoded_text` on label sequences (single-char tokens)."""
    n, k = len(labels), len(keyword_labels)
    if k == 0:
        return True
    return any(list(labels[i:i + k]) == list(keyword_labels) for i in range(n - k + 1))


def detect_confidence(labels, keyword_labels):
    """calculate_confidence, ml_models/test.py:230-235: 0.9 on a hit else 0."""
    return 0.9 if keyword_hit(labels, keyword_labels) else 0.0


# ----------------------------------------------------------------------------
# CTC loss
# ----------------------------------------------------------------------------
def ctc_loss_torch(log_probs, targets, input_lengths, target_lengths, blank=0,
                   reduction="mean", zero_infinity=False, want_grad=True, dtype="float32"):
    """torch.nn.functional.ctc_loss on CPU.  log_probs [T,B,C] float32.

    Returns (loss, grad wrt log_probs or None).
    """
    import torch
    import torch.nn.functional as F

    lp = torch.tensor(np.asarray(log_probs), dtype=torch.float64 if dtype == "float64" else torch.float32,
                      requires_grad=want_grad)
    loss = F.ctc_loss(lp, torch.as_tensor(np.asarray(targets)).long(),
                      torch.as_tensor(np.asarray(input_lengths)).long(),
                      torch.as_tensor(np.asarray(target_lengths)).long(),
                      blank=blank, reduction=reduction, zero_infinity=zero_infinity)
    grad = None
    if want_grad:
        (loss.sum() if reduction == "none" else loss).backward()
        grad = lp.grad.numpy()
    return loss.detach().numpy(), grad


def _logaddexp(a, b):
    return np.logaddexp(a, b)


def ctc_loss_numpy64(log_probs, targets, input_lengths, target_lengths, blank=0):
    """Independent alpha/beta CTC in fp64.  Returns (nll[B], grad[T,B,C]).

    grad follows PyTorch's convention (native ctc_loss_backward): the gradient
    of the per-sample nll with respect to the UN-NORMALISED activations whose
    log-softmax is log_probs:  exp(lp) - exp(log sum_{s: l'_s=c} alpha beta + nll - lp),
    zero for t >= input_length.
    """
    lp = np.asarray(log_probs, dtype=np.float64)
    T, B, C = lp.shape
    nll = np.zeros(B)
    grad = np.zeros_like(lp)
    NEG = -np.inf
    for b in range(B):
        Tb = int(input_lengths[b])
        S = int(target_lengths[b])
        tgt = [int(v) for v in np.asarray(targets)[b][:S]]
        ext = [blank]
        for v in tgt:
            ext += [v, blank]
        L = len(ext)
        la = np.full((Tb, L), NEG)
        lb = np.full((Tb, L), NEG)
        if Tb == 0:
            nll[b] = 0.0 if S == 0 else np.inf
            continue
        la[0, 0] = lp[0, b, blank]
        if L > 1:
            la[0, 1] = lp[0, b, ext[1]]
        for t in range(1, Tb):
            for s in range(L):
                v = la[t - 1, s]
                if s >= 1:
                    v = _logaddexp(v, la[t - 1, s - 1])
                if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                    v = _logaddexp(v, la[t - 1, s - 2])
                la[t, s] = v + lp[t, b, ext[s]]
        ll = la[Tb - 1, L - 1]
        if L > 1:
            ll = _logaddexp(ll, la[Tb - 1, L - 2])
        nll[b] = -ll
        lb[Tb - 1, L - 1] = lp[Tb - 1, b, blank]
        if L > 1:
            lb[Tb - 1, L - 2] = lp[Tb - 1, b, ext[L - 2]]
        for t in range(Tb - 2, -1, -1):
            for s in range(L):
                v = lb[t + 1, s]
                if s + 1 < L:
                    v = _logaddexp(v, lb[t + 1, s + 1])
                if s + 2 < L and ext[s] != blank and ext[s] != ext[s + 2]:
                    v = _logaddexp(v, lb[t + 1, s + 2])
                lb[t, s] = v + lp[t, b, ext[s]]
        for t in range(Tb):
            lab = np.full(C, NEG)
            for s in range(L):
                lab[ext[s]] = _logaddexp(lab[ext[s]], la[t, s] + lb[t, s])
            with np.errstate(invalid="ignore", over="ignore"):
                grad[t, b] = np.exp(lp[t, b]) - np.exp(lab + nll[b] - lp[t, b])
    return nll, grad
