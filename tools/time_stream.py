#!/usr/bin/env python
"""A/B timing of the streaming config (one 1-hour stream, python CMVN, tensor CNN) and the fused clip path;
WW_B200_LIB picks the library build."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


dev = torch.device("cuda", 0)
sd = bench.load_weights()
n = 3600 * 16000
pcm = torch.randint(-2000, 2000, (n,), dtype=torch.int16, device=dev)
sc = ww_b200.StreamScorer(sd, device=0, cmvn="python", cnn_impl="tensor")
dt = timed(lambda: sc.score(pcm))
clips = bench.synth_pcm(262144, dev, 1234)
ws = ww_b200.WakeWordScorer(sd, device=0, cnn_impl="tensor")
dt2 = timed(lambda: ws.score(clips), reps=5)
print("%s: stream %.3f ms (%.1f M windows/s), fused %.2f M clips/s" % (os.environ.get("WW_B200_LIB", "default"), dt * 1e3, 224939 / dt / 1e6, 262144 / dt2 / 1e6))
