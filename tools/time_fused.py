#!/usr/bin/env python
"""A/B timing of ww_score_clips over N resident clips: chunked launches (WW_OPT_FUSED = 0) against the one-kernel path
(WW_OPT_FUSED = 2) for several splits of the SMs between the CNN role and the frontend pipelines.
Usage: python tools/time_fused.py [clips (262144)] [cmvn python|device] [cnn SM counts, comma separated]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402
from ww_b200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cmvn = sys.argv[2] if len(sys.argv) > 2 else "python"
sms = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 9, 10, 11, 12]
dev = torch.device("cuda", 0)
pcm = bench.synth_pcm(n, dev, 1234)
sd = dict(np.load(os.path.join(ROOT, "tests", "golden", "xiaoa_weights.npz")))
sc = ww_b200.WakeWordScorer(sd, device=0, cmvn=cmvn, decision="python" if cmvn == "python" else "device")
ctx = sc.ctx


def run(tag):
    for _ in range(2):
        out = sc.score(pcm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(5):
            out = sc.score(pcm)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print("%-28s %8.3f ms  %7.2f M clips/s" % (tag, best, n / best / 1e3), flush=True)
    return out


ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_FUSED, 0), "opt")
ref = run("chunked launches")
ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_FUSED, 2), "opt")
for k in sms:
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_FUSED_CNN_SMS, k), "opt")
    out = run("one kernel, %2d CNN SMs" % k)
    same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    print("    identical to the chunked path: %s" % same, flush=True)
f = ww_b200.mfcc_batch(pcm[: min(n, 262144)])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f = ww_b200.mfcc_batch(pcm[: min(n, 262144)])
e1.record()
torch.cuda.synchronize()
print("frontend alone: %.2f M clips/s" % (min(n, 262144) / (e0.elapsed_time(e1) / 5) / 1e3))
if hasattr(ctx.lib, "ww_debug_fused_stats"):
    # experiment build (-DWW_FUSED_STATS): where the two roles wait
    import ctypes as C
    out = (C.c_ulonglong * (8 + 1024))()
    ctx.lib.ww_debug_fused_stats(ctx.h, out)
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_FUSED_CNN_SMS, 10), "opt")
    sc.score(pcm)
    ctx.lib.ww_debug_fused_stats(ctx.h, out)
    pw, cw, ct, pt = [int(v) for v in out][:4]
    print('producer slow-path waits: %d, longest %d cycles, last waiting octet %d, waits > 20000 cycles: %d' % tuple(int(v) for v in out[4:8]))
    print("stats (one call, 10 CNN SMs): producer DCT-warp wait %.3e of pipeline-warp cycles %.3e (4 of 8 warps wait: %.1f %% of "
          "their time); consumer wait %.3e of %.3e warp cycles (%.1f %%)" % (pw, pt, 100.0 * pw / (pt / 2), cw, ct, 100.0 * cw / ct))
    per = np.array([int(v) for v in out[8:8 + 296]], dtype=np.float64).reshape(148, 2)
    wt = np.array([int(v) for v in out[8 + 512:8 + 512 + 296]], dtype=np.float64).reshape(148, 2) / 4
    live = per[:, 0] > 0
    print("pipelines: %d, total cycles min %.3e avg %.3e max %.3e; busy (total - wait) min %.3e avg %.3e max %.3e" % (
        2 * live.sum(), per[live].min(), per[live].mean(), per[live].max(), (per - wt)[live].min(), (per - wt)[live].mean(), (per - wt)[live].max()))
    busy = (per - wt)[live].mean(axis=1)
    order = np.argsort(busy)
    sm_ids = np.nonzero(live)[0]
    print("slowest CTAs (busy cycles):", [(int(sm_ids[i]), "%.3e" % busy[i]) for i in order[-8:]])
    print("fastest CTAs (busy cycles):", [(int(sm_ids[i]), "%.3e" % busy[i]) for i in order[:8]])
