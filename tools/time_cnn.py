#!/usr/bin/env python
"""CNN kernels alone: windows/s of ww_cnn_forward over N resident [13,63] windows (CUDA events, best of 3 x 10 launches)
for the tcgen05 fp16 kernel (with its guard-band re-score), the int8 twin and the exact fp32 kernel; flat batches and the
sliding windows of a stream.  WW_B200_LIB picks the library build (A/B)."""
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
from ww_b200.model import XIAOA_EXPONENTS, _push_weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
sd = bench.load_weights()
ctx = L.get_context(0)
_push_weights(ctx, sd, ("time-cnn", 0), XIAOA_EXPONENTS)
g = torch.Generator(device=dev)
g.manual_seed(1)
feats = torch.randn((n, 13, 63), generator=g, device=dev) * 8 - 10
logits = torch.empty((n, 1), device=dev)
dec = torch.empty((n,), dtype=torch.uint8, device=dev)
sp = L.cur_stream(dev)
tag = os.path.basename(os.environ.get("WW_B200_LIB", "default"))


def run(impl, cmvn, decide, thr):
    def f():
        ctx.check(ctx.lib.ww_cnn_forward(ctx.h, L.ptr(feats), 819, 63, 1, n, cmvn, decide, thr, impl, L.ptr(logits),
                                         L.ptr(dec), sp), "fwd")
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    return best


for name, impl, cmvn, decide, thr in (("tensor fp16, python CMVN", L.CNN_TENSOR, L.CMVN_PY, L.DECIDE_LOGIT, 0.0),
                                      ("tensor fp16, device CMVN", L.CNN_TENSOR, L.CMVN_DEVICE, L.DECIDE_DEVICE, 80.0),
                                      ("int8 twin, device CMVN", L.CNN_INT8, L.CMVN_DEVICE, L.DECIDE_DEVICE, 80.0)):
    ms = run(impl, cmvn, decide, thr)
    print(f"{tag}: {name:28s} {n} windows  {ms:8.3f} ms  {n / ms / 1e3:8.1f} M windows/s  "
          f"{n * 1291968 / ms / 1e9:7.1f} TFLOP/s useful")
if n <= 1 << 18:
    ms = run(L.CNN_FP32, L.CMVN_PY, L.DECIDE_LOGIT, 0.0)
    print(f"{tag}: {'exact fp32 (CUDA cores)':28s} {n} windows  {ms:8.3f} ms  {n / ms / 1e3:8.1f} M windows/s")
