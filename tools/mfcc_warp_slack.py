#!/usr/bin/env python
"""Per-warp slack at the block barrier of the clip-shape frontend (diagnostic build: nvcc ... -DWW_MFCC_STATS -o
build_ab/lib_stats.so; WW_B200_LIB=build_ab/lib_stats.so python tools/mfcc_warp_slack.py [clips]).
Prints, per warp index, the mean clocks between the warp's arrival at the barrier and the barrier's completion --
the warp with the SMALLEST number is the one the CTA waits for."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "esp32-wake-word_b200")]
import bench  # noqa: E402
import ww_b200  # noqa: E402
from ww_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
pcm = bench.synth_pcm(n, dev, 1234)
lib = _lib.load_library()
lib.ww_debug_mfcc_slack.argtypes = [C.c_void_p, C.c_int]
buf = np.zeros(296 * 8, np.uint64)
f = ww_b200.mfcc_batch(pcm)
torch.cuda.synchronize()
lib.ww_debug_mfcc_slack(buf.ctypes.data, buf.size)   # clear
f = ww_b200.mfcc_batch(pcm)
torch.cuda.synchronize()
lib.ww_debug_mfcc_slack(buf.ctypes.data, buf.size)
blocks_per_cta = 2 * n / 296
s = buf.reshape(296, 8).astype(np.float64) / blocks_per_cta
print("mean clocks from arrival to barrier completion, per warp (all CTAs):", np.round(s.mean(0), 1))
print("even CTAs (block 0 of a clip: frame 0 is the edge frame):", np.round(s[0::2].mean(0), 1))
print("odd CTAs  (block 1: tail frame, 31 valid frames):        ", np.round(s[1::2].mean(0), 1))
print("per-CTA mean over warps: min %.1f max %.1f" % (s.mean(1).min(), s.mean(1).max()))
