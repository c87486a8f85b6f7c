#!/usr/bin/env python
"""ncu driver for the int8 tensor-core kernel: a few launches of ww_cnn_forward_i8 over N random windows."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = torch.randint(-128, 128, (n, 13, 63), dtype=torch.int8, device="cuda")
sd = bench.load_weights()
for _ in range(3):
    out = ww_b200.forward_int8(sd, x)
    torch.cuda.synchronize()
print("done", n)
