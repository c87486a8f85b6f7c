"""Push/poll session timing for one build of the library (A/B: WW_B200_LIB=<path> python tools/time_session.py).

4096 concurrent streams, one 20 ms chunk (320 samples) per push from pinned host memory, hits polled after every
push -- the measurement of tools/bench_configs.py:session_bench, repeated so that run-to-run spread is visible."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "esp32-wake-word_b200")]
import ww_b200  # noqa: E402
from ww_b200 import _lib  # noqa: E402


def main():
    n_streams, chunk, pushes = 4096, 320, 300
    d = np.load(os.path.join(ROOT, "tests", "golden", "xiaoa_weights.npz"))
    sd = {k: d[k] for k in d.files}
    rng = np.random.default_rng(3)
    data = torch.from_numpy((rng.standard_normal((n_streams, chunk)) * 600).astype(np.int16)).pin_memory()
    for cmvn, impl in (("device", "tensor"), ("python", "tensor")):
        best = []
        for rep in range(3):
            ses = ww_b200.StreamSession(sd, n_streams, max_chunk_samples=chunk, device=0, cmvn=cmvn, cnn_impl=impl)
            for _ in range(70):
                ses.write(data)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(pushes):
                ses.write(data)
                ses.poll()
            best.append((time.perf_counter() - t0) / pushes * 1e3)
            ses.close()
        print(f"{os.path.basename(_lib.LIB_PATH)} cmvn={cmvn}: ms per push {' '.join(f'{b:.3f}' for b in best)}"
              f" -> {n_streams * (chunk / 16000) / (min(best) * 1e-3) / 1e3:.0f} k real-time streams", flush=True)


if __name__ == "__main__":
    main()
