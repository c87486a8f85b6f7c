"""GPU parity: sliding-window stream scoring against the oracle (esp_wake_word_detector.cpp semantics)."""
import numpy as np
import pytest
import torch

from oracle import mfcc as om
from oracle import stream as ostream

pytestmark = pytest.mark.gpu


def _stream(seconds, seed):
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    x = rng.normal(0, 0.02, n)
    for k in range(0, int(seconds), 3):  # a 1 s burst every 3 s
        s = k * 16000 + 4000
        x[s:s + 16000] += rng.normal(0, 0.2, min(16000, n - s))
    return np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("cmvn", ["python", "device"])
def test_stream_windows_vs_oracle(cuda_device, xiaoa_sd, cmvn):
    import ww_b200

    pcm = _stream(20.0, seed=4321)
    feats_o, logits_o = ostream.window_logits(om.pcm16_to_float(pcm), xiaoa_sd, cmvn=cmvn)
    sc = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn=cmvn, cnn_impl="fp32")
    feats, logits = sc.score(torch.from_numpy(pcm).to(cuda_device))
    torch.cuda.synchronize()
    T = 1 + len(pcm) // 256
    assert feats.shape == (13, T) and logits.shape == (T - 62, 1)
    assert np.abs(feats.cpu().numpy() - feats_o).max() < 1e-3
    got = logits.cpu().numpy()
    if cmvn == "python":
        assert np.abs(got - logits_o).max() < 2e-3
    else:
        assert np.mean(np.abs(got - logits_o) > 2e-3) < 0.03  # int8 rounding flips, see test_gpu_cnn
    ev_o = ostream.events(logits_o)
    ev = ww_b200.events(got)
    clear = np.abs(logits_o[:, 0] - ostream.LN4) > 1e-2
    if clear.all():
        assert ev == ev_o


def test_stream_equals_clip_scoring_on_aligned_windows(cuda_device, xiaoa_sd):
    """Window w of a stream (frames w..w+62) must see the same features as the fused clip scorer would
    compute for... itself: the stream path and the [B,13,63] path share the CNN kernel; check stride handling."""
    import ww_b200
    from ww_b200 import _lib as L

    pcm = _stream(6.0, seed=1)
    sc = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn="python", cnn_impl="fp32")
    feats, logits = sc.score(torch.from_numpy(pcm).to(cuda_device))
    W = logits.shape[0]
    win = torch.stack([feats[:, w:w + 63] for w in range(0, W, 7)])  # explicit [n,13,63] copies
    m = ww_b200.LightweightKWS(1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in xiaoa_sd.items()})
    m.cnn_impl = "fp32"
    out = m(ww_b200.cmvn_batch(win))
    torch.cuda.synchronize()
    assert torch.allclose(out[:, 0], logits[0:W:7, 0], atol=1e-5)
