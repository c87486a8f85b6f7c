#!/usr/bin/env python
"""Randomised consistency sweep on the GPU box (not part of the test-suite): odd batch sizes, ragged lengths, random
chunkings -- every path is compared with another path of the library that must give identical results."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
dev = torch.device("cuda", 0)
sd = bench.load_weights()
bad = 0


def check(name, ok):
    global bad
    if not ok:
        bad += 1
        print("MISMATCH", name, flush=True)


# 1. frontend: a clip scores the same alone and inside any batch, int16 and fp32 PCM agree to 1e-4
base = bench.synth_pcm(257, dev, 11)
f_all = ww_b200.mfcc_batch(base)
for n in rng.integers(1, 257, size=6):
    idx = torch.from_numpy(rng.choice(257, size=int(n), replace=False)).to(dev)
    check(f"mfcc batch-invariance n={n}", torch.equal(ww_b200.mfcc_batch(base[idx]), f_all[idx]))
check("mfcc fp32 vs int16", (ww_b200.mfcc_batch(base.float() / 32768.0) - f_all).abs().max().item() < 1e-3)
# ragged lengths: features of a prefix equal the first frames of a zero-padded... (only frames not touching the end)
for L in rng.integers(600, 16000, size=5):
    L = int(L) // 8 * 8
    f = ww_b200.mfcc_batch(base[:9, :L].contiguous())
    T = f.shape[2]
    safe = max(0, (L - 416 + 256) // 256 - 1)
    check(f"mfcc prefix L={L}", torch.equal(f[:, :, :safe], f_all[:9, :, :safe]))

# 2. CNN: tensor path decisions == fp32 path decisions, logits within 1e-2; odd window counts
feats = f_all
for n in list(rng.integers(1, 257, size=5)) + [8, 31, 33]:
    n = int(n)
    a = ww_b200.WakeWordScorer(sd, device=0, cnn_impl="fp32")
    l0, d0 = a.score(base[:n])
    b = ww_b200.WakeWordScorer(sd, device=0, cnn_impl="tensor")
    l1, d1 = b.score(base[:n])
    check(f"tensor decisions n={n}", torch.equal(d0, d1))
    check(f"tensor logits n={n}", (l0 - l1).abs().max().item() < 1e-2)
    c = ww_b200.WakeWordScorer(sd, device=0, cmvn="device", decision="device", cnn_impl="int8")
    l2, d2 = c.score(base[:n])
    oq, dq = ww_b200.score_clips_int8(sd, base[:n])
    check(f"int8 scorer n={n}", torch.equal(l2 * 8.0, oq.float()) and torch.equal(d2, dq))

# 3. sessions: random chunkings reproduce whole-stream scoring, all three CNN implementations
stream = bench.synth_pcm(5, dev, 23).reshape(-1).cpu().numpy()
for impl, cmvn in (("tensor", "python"), ("fp32", "device"), ("int8", "device")):
    ss = ww_b200.StreamScorer(sd, device=0, cmvn=cmvn, cnn_impl=impl)
    _, whole = ss.score(torch.from_numpy(stream).to(dev))
    whole = whole.cpu().numpy()
    ses = ww_b200.StreamSession(sd, 1, max_chunk_samples=8000, cmvn=cmvn, cnn_impl=impl)
    pos, got = 0, []
    while pos < len(stream):
        n = int(rng.integers(1, 1000)) * 8
        n = min(n, len(stream) - pos)
        n -= n % 8
        if n == 0:
            break
        got.append(ses.write(stream[None, pos:pos + n]))
        pos += n
    ses.close()
    g = np.concatenate([x for x in got if x.shape[1]], axis=1)[0]
    if impl == "tensor":
        check(f"session {impl}", np.abs(g - whole[:len(g)]).max() < 1e-2)
    else:
        check(f"session {impl}", np.array_equal(g, whole[:len(g)]))

# 4. CTC: greedy two-phase path == torch argmax + collapse on random shapes
for _ in range(6):
    B, T, C = int(rng.integers(1, 9)), int(rng.integers(1, 300)), int(rng.choice([33, 40, 64, 100, 257, 1000]))
    lp = torch.randn(B, T, C, device=dev)
    lab, n, _ = ww_b200.greedy_batch(lp, mode="collapse")
    am = lp.argmax(-1).cpu().numpy()
    for b in range(B):
        want, prev = [], 0
        for t in range(T):
            if am[b, t] != 0 and am[b, t] != prev:
                want.append(int(am[b, t]))
            prev = am[b, t]
        check(f"greedy B={B} T={T} C={C}", lab[b, :int(n[b])].tolist() == want)

# 5. TDM down-mix: random shapes against the integer formula in torch
for _ in range(5):
    B, n = int(rng.integers(1, 5)), int(rng.integers(1, 5000))
    x = torch.randint(-32768, 32768, (B, 12 * n), dtype=torch.int16, device=dev)
    fr = x.view(B, 3 * n, 4).to(torch.int32)
    w = (fr[..., 0] << 6) + (fr[..., 1] << 5) + (fr[..., 2] << 6)
    mono = (w >> 7).to(torch.int16).to(torch.int32).view(B, n, 3)
    want = ((mono[..., 0] + 2 * mono[..., 1] + mono[..., 2]) >> 2).to(torch.int16)
    check(f"tdm B={B} n={n}", torch.equal(ww_b200.tdm_downmix(x), want))

torch.cuda.synchronize()
print("fuzz done, mismatches:", bad)
sys.exit(1 if bad else 0)
