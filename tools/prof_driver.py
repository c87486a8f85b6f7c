#!/usr/bin/env python
"""Small driver for ncu: a few launches of each hot kernel on synthetic clips (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
what = sys.argv[2] if len(sys.argv) > 2 else "all"
impl = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda", 0)
pcm = bench.synth_pcm(n, dev, 1234)
sd = bench.load_weights()
sc = ww_b200.WakeWordScorer(sd, device=0, cnn_impl=impl)
torch.cuda.synchronize()
for _ in range(3):
    if what in ("all", "mfcc"):
        f = ww_b200.mfcc_batch(pcm)
    if what in ("all", "fused"):
        lg, dec = sc.score(pcm)
    torch.cuda.synchronize()
print("done", n, what)
