#!/usr/bin/env python
"""Pinned host -> device copy ceiling of the box (the bound of bench.py's `e2e`): 54.4-55.5 GB/s measured on the B200
pool, i.e. 1.70-1.73 M one-second int16 clips per second per GPU."""
import torch,time
x=torch.empty(1<<31,dtype=torch.uint8).pin_memory()
d=torch.empty(1<<31,dtype=torch.uint8,device='cuda')
for n in (1<<31,1<<28,1<<25):
    torch.cuda.synchronize()
    for _ in range(2): d[:n].copy_(x[:n],non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    reps=max(1,(1<<32)//n)
    for _ in range(reps): d[:n].copy_(x[:n],non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(n, "bytes/copy:", reps*n/dt/1e9, "GB/s")
