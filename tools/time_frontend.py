#!/usr/bin/env python
"""A/B timing of the frontend kernel: clips/s of ww_mfcc_batch over N resident clips (CUDA events, 10 launches),
plus the max deviation from a reference output file if given.  WW_B200_LIB picks the library build."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
pcm = bench.synth_pcm(n, dev, 1234)
for _ in range(3):
    f = ww_b200.mfcc_batch(pcm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record()
    for _ in range(10):
        f = ww_b200.mfcc_batch(pcm)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
ref_path = os.path.join(ROOT, "gpurun_out", "ab_ref.npy")
sub = f[:4096].cpu().numpy()
dev_str = ""
if os.path.exists(ref_path):
    dev_str = " max|d| vs first variant %.3e" % float(np.abs(sub - np.load(ref_path)).max())
else:
    os.makedirs(os.path.dirname(ref_path), exist_ok=True)
    np.save(ref_path, sub)
print("%s: %.3f ms per launch, %.2f M clips/s%s" % (os.environ.get("WW_B200_LIB", "default"), best, n / best / 1e3, dev_str))
