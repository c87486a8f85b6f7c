#!/usr/bin/env python
"""One CTC loss fwd+bwd at the ctc.py-like shape (T=801, B=256, C=4096, S=32) -- for `ncu` launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import ww_b200  # noqa: E402

T, B, C, S = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (801, 256, 4096, 32)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(777)
lp = torch.log_softmax(torch.randn((T, B, C), generator=g, device=dev), dim=-1)
tg = torch.randint(1, C, (B, S), generator=g, device=dev)
il = torch.full((B,), T, dtype=torch.int32, device=dev)
tl = torch.full((B,), S, dtype=torch.int32, device=dev)
crit = ww_b200.CTCLoss(blank=0, zero_infinity=True)
for _ in range(3):
    x = lp.detach().requires_grad_(True)
    crit(x, tg, il, tl).backward()
    torch.cuda.synchronize()
print("done")
