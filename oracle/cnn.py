"""LightweightKWS oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference restated: ml_models/src/wakeModel.py:4-34
    conv_layers: 3 x [Conv1d(k=3, padding=1, bias=False) -> ReLU -> MaxPool1d(2)]
                 channels 13 -> 32 -> 64 -> 128   (T: 63 -> 31 -> 15 -> 7)
    global_pool: AdaptiveAvgPool1d(1)
    classifier : Linear(128, 64, bias=False) -> ReLU -> Linear(64, C, bias=False)
Decision: torch.sigmoid(out) > 0.5 (ml_models/main.py:53,111,123) and the
device form sigmoid*100 >= 80 (esp_wake_word_detector.cpp:226-228,245).

Pinning: the int8 power-of-two twin below reproduces the only known-answer
vector the reference ships (ml_models/xiaoa.info:3153-3224, int8 input
[1,63,13] at exponent -4 -> int8 output -40 at exponent -3 = -5.0) exactly;
the fp32 forward is additionally pinned against the reference's own
`wakeModel.LightweightKWS` imported in the build container
(tests/golden/make_golden.py -> tests/golden/kws_forward.npz).
"""
from __future__ import annotations

import numpy as np


def forward_torch(x, sd, dtype="float32"):
    """x: [B, 13, T] -> logits [B, C] with torch CPU ops (the reference path)."""
    import torch
    import torch.nn.functional as F

    dt = torch.float64 if dtype == "float64" else torch.float32
    x = torch.as_tensor(np.asarray(x)).to(dt)
    w = {k: torch.as_tensor(np.asarray(v)).to(dt) for k, v in sd.items()}
    with torch.no_grad():
        for key in ("conv_layers.0.weight", "conv_layers.3.weight", "conv_layers.6.weight"):
            x = F.max_pool1d(F.relu(F.conv1d(x, w[key], padding=1)), 2)
        x = x.mean(dim=-1)  # AdaptiveAvgPool1d(1) + squeeze
        x = F.relu(x @ w["classifier.0.weight"].T)
        x = x @ w["classifier.2.weight"].T
    return x.numpy()


def forward_numpy64(x, sd):
    """Index-level fp64 forward (no torch): x [B,13,T] -> [B,C]."""
    a = np.asarray(x, dtype=np.float64)
    for key in ("conv_layers.0.weight", "conv_layers.3.weight", "conv_layers.6.weight"):
        w = np.asarray(sd[key], dtype=np.float64)  # [O, I, 3]
        B, I, T = a.shape
        ap = np.pad(a, ((0, 0), (0, 0), (1, 1)))
        y = np.zeros((B, w.shape[0], T))
        for r in range(3):
            y += np.einsum("oi,bit->bot", w[:, :, r], ap[:, :, r:r + T])
        y = np.maximum(y, 0.0)
        Tp = T // 2
        a = np.maximum(y[:, :, 0:2 * Tp:2], y[:, :, 1:2 * Tp:2])
    g = a.mean(axis=-1)
    h = np.maximum(g @ np.asarray(sd["classifier.0.weight"], dtype=np.float64).T, 0.0)
    return h @ np.asarray(sd["classifier.2.weight"], dtype=np.float64).T


def decide_python(logits):
    """torch.sigmoid(out) > 0.5  <=>  logit > 0 (ml_models/main.py:53)."""
    return np.asarray(logits) > 0.0


def decide_device(logits):
    """1/(1+expf(-x))*100 >= 80 (esp_wake_word_detector.cpp:226-228,245)."""
    x = np.asarray(logits, dtype=np.float32)
    s = np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32)) * np.float32(100.0)
    return s >= np.float32(80.0)


# ----------------------------------------------------------------------------
# int8 power-of-two twin (esp-dl / esp_ppq export), SURVEY.md Appendix B
# ----------------------------------------------------------------------------
EXPONENTS = {
    "input": -4,
    "w1": -8, "a1": -5,
    "w2": -9, "a2": -5,
    "w3": -9, "a3": -4,
    "gap": -5,
    "wf1": -9, "f1": -4,
    "wf2": -9, "out": -3,
}


def _round_half_up(x):
    return np.floor(np.asarray(x, dtype=np.float64) + 0.5)


def _round_half_even(x):
    return np.rint(np.asarray(x, dtype=np.float64))


def quantize(x, exponent, rounding=_round_half_even):
    return np.clip(rounding(np.asarray(x, dtype=np.float64) / 2.0 ** exponent), -128, 127).astype(np.int64)


def forward_int8(x_q, sd, exps=EXPONENTS, rounding=_round_half_even):
    """x_q: int8 input [B, 13, T] at exponent exps['input'] -> int8 output [B, C].

    Every tensor is symmetric per-tensor power-of-two int8 (xiaoa.json:5-20).
    Integer accumulation, then requantisation by a right shift with `rounding`.
    """
    def q_w(key, e):
        return quantize(sd[key], e)

    def requant(acc, e_acc, e_out):
        return np.clip(rounding(acc.astype(np.float64) * 2.0 ** (e_acc - e_out)), -128, 127).astype(np.int64)

    a = np.asarray(x_q, dtype=np.int64)
    e_a = exps["input"]
    for key, we, ae in (("conv_layers.0.weight", "w1", "a1"),
                        ("conv_layers.3.weight", "w2", "a2"),
                        ("conv_layers.6.weight", "w3", "a3")):
        w = q_w(key, exps[we])
        B, I, T = a.shape
        ap = np.pad(a, ((0, 0), (0, 0), (1, 1)))
        acc = np.zeros((B, w.shape[0], T), dtype=np.int64)
        for r in range(3):
            acc += np.einsum("oi,bit->bot", w[:, :, r], ap[:, :, r:r + T])
        acc = np.maximum(acc, 0)  # ReLU fused before requantisation
        y = requant(acc, e_a + exps[we], exps[ae])
        Tp = T // 2
        a = np.maximum(y[:, :, 0:2 * Tp:2], y[:, :, 1:2 * Tp:2])
        e_a = exps[ae]
    # global average pool: mean over T in real arithmetic, requantised
    g = a.sum(axis=-1).astype(np.float64) / a.shape[-1] * 2.0 ** e_a
    g = quantize(g, exps["gap"], rounding)
    w = q_w("classifier.0.weight", exps["wf1"])
    acc = np.maximum(g @ w.T, 0)
    h = requant(acc, exps["gap"] + exps["wf1"], exps["f1"])
    w = q_w("classifier.2.weight", exps["wf2"])
    acc = h @ w.T
    return requant(acc, exps["f1"] + exps["wf2"], exps["out"])
