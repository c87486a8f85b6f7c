#!/usr/bin/env python
"""Host-only timing of ww_wav_load_batch: N one-second 16-bit mono WAV files from the page cache
into one int16 batch, files/s for a few thread counts.  No GPU needed.  WW_B200_LIB picks the library build."""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "esp32-wake-word_b200")]
from ww_b200 import _lib as L  # noqa: E402
from ww_b200 import wav  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    lib = L.load_library()
    rng = np.random.default_rng(0)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        clips = rng.integers(-3000, 3000, (64, 16000)).astype(np.int16)
        paths = []
        for i in range(n):
            p = os.path.join(d, f"c{i:06d}.wav")
            wav.write_wav(p, clips[i % 64])
            paths.append(p)
        arr = (C.c_char_p * n)(*[p.encode() for p in paths])
        pcm = np.zeros((n, 16000), np.int16)
        status = (C.c_int * n)()
        for threads in (1, 4, len(os.sched_getaffinity(0))):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                failed = lib.ww_wav_load_batch(arr, n, 16000, threads, pcm.ctypes.data_as(C.c_void_p), None, status)
                best = min(best, time.perf_counter() - t0)
            assert failed == 0 and np.array_equal(pcm[:64], clips) and np.array_equal(pcm[n - 1], clips[(n - 1) % 64])
            print(f"{os.path.basename(L.LIB_PATH)} threads={threads}: {n / best / 1e3:.0f} k files/s "
                  f"({n * 32044 / best / 1e9:.2f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
