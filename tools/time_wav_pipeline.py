#!/usr/bin/env python
"""files -> decisions on the GPU box: N one-second 16-bit WAV files (page cache, /dev/shm) through
  (a) load_wav_batch, then score_host          (round 1's score_wav_dir: load everything, then score)
  (b) score_wav_files                          (two-buffer pipeline inside the library)
files/s for both, with the library's own split of the pipeline's time (reading / blocked on the GPU)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "esp32-wake-word_b200")]
import bench  # noqa: E402
import ww_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
threads = len(os.sched_getaffinity(0))
rng = np.random.default_rng(0)
sd = bench.load_weights()
sc = ww_b200.WakeWordScorer(sd, device=0)
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
with tempfile.TemporaryDirectory(dir=base) as d:
    clips = np.clip(np.round(rng.normal(0, 0.1, (256, 16000)) * 32767), -32768, 32767).astype(np.int16)
    paths = []
    for i in range(n):
        p = os.path.join(d, f"c{i:06d}.wav")
        ww_b200.write_wav(p, clips[i % 256])
        paths.append(p)
    for rep in range(2):
        t0 = time.perf_counter()
        pcm, _, _ = ww_b200.load_wav_batch(paths, threads=threads)
        t1 = time.perf_counter()
        la, da = sc.score_host(pcm)
        t2 = time.perf_counter()
        lb, db, _, _, stats = ww_b200.score_wav_files(paths, sc, threads=threads)
        t3 = time.perf_counter()
        assert np.array_equal(la, lb) and np.array_equal(da, db)
        print(json.dumps({"config": "files -> decisions, one-second 16-bit WAV files from the page cache", "files": n,
                          "threads": threads, "pass": rep,
                          "load_then_score": {"load_s": t1 - t0, "score_host_s": t2 - t1, "files_per_s": n / (t2 - t0)},
                          "pipelined": {"total_s": t3 - t2, "files_per_s": n / (t3 - t2), "library_stats": stats},
                          "speedup": (t2 - t0) / (t3 - t2)}), flush=True)
        del pcm
