"""GPU parity: push/poll streaming sessions equal whole-stream scoring, for any chunking and many streams."""
import numpy as np
import pytest
import torch

from oracle import stream as ostream

pytestmark = pytest.mark.gpu


def _streams(n_streams, seconds, seed):
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    x = rng.normal(0, 0.03, (n_streams, n))
    for s in range(n_streams):
        for k in range(s % 3, int(seconds), 4):
            a = k * 16000 + 3000
            x[s, a:a + 16000] += rng.normal(0, 0.25, min(16000, n - a))
    return np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("chunk", [16000, 4096, 1600, 320])
@pytest.mark.parametrize("cmvn,impl", [("device", "fp32"), ("python", "tensor")])
def test_session_equals_whole_stream(cuda_device, xiaoa_sd, chunk, cmvn, impl):
    import ww_b200

    S, seconds = 5, 12
    pcm = _streams(S, seconds, seed=chunk)
    n = pcm.shape[1] // chunk * chunk
    pcm = pcm[:, :n]
    # the reference is always the exact kernel: a tensor-path session re-scores the windows inside the guard band of ITS
    # threshold in fp32, so its hit list must be the fp32 path's hit list
    ref = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn=cmvn, cnn_impl="fp32")
    want = [ref.score(torch.from_numpy(pcm[s]).to(cuda_device))[1].cpu().numpy() for s in range(S)]

    thr = float(np.percentile(np.concatenate(want), 97))   # make a few hits happen
    sess = ww_b200.StreamSession(xiaoa_sd, S, max_chunk_samples=16000, device=0, cmvn=cmvn, cnn_impl=impl,
                                 threshold_logit=thr, refractory=40)
    got = [[] for _ in range(S)]
    hits = []
    for c0 in range(0, n, chunk):
        lg = sess.write(pcm[:, c0:c0 + chunk])
        for s in range(S):
            got[s].append(lg[s])
        hits += sess.poll()
    t_count = (n - 160) // 256 + 1            # frames whose taps are complete (the stream end is never padded)
    for s in range(S):
        g = np.concatenate(got[s], axis=0)
        assert g.shape[0] == t_count - 62 == sess.windows
        w = want[s][: g.shape[0]]
        if impl == "fp32":
            np.testing.assert_array_equal(g, w)           # same kernels, same inputs: bit-identical
        else:
            assert np.abs(g - w).max() < 1e-2
        want_hits = ostream.events(w, threshold_logit=thr, refractory=40)
        assert [h[1] for h in hits if h[0] == s] == want_hits
    assert sess.poll() == []
    sess.close()


def test_session_rejects_bad_chunks(cuda_device, xiaoa_sd):
    import ww_b200

    sess = ww_b200.StreamSession(xiaoa_sd, 2, max_chunk_samples=800, device=0)
    with pytest.raises(ww_b200.WWError):
        sess.write(np.zeros((2, 804), np.int16))      # larger than max_chunk_samples
    with pytest.raises(ww_b200.WWError):
        sess.write(np.zeros((2, 100), np.int16))      # not a multiple of 8
    with pytest.raises(ValueError):
        sess.write(np.zeros((3, 800), np.int16))      # wrong stream count
    assert sess.write(np.zeros((2, 800), np.int16)).shape == (2, 0, 1)


def test_session_many_streams(cuda_device, xiaoa_sd):
    """2048 concurrent streams, 20 ms pushes for 2.5 s: one launch pair per push."""
    import ww_b200

    S = 2048
    rng = np.random.default_rng(0)
    sess = ww_b200.StreamSession(xiaoa_sd, S, max_chunk_samples=320, device=0, cmvn="device", cnn_impl="tensor")
    total = 0
    base = np.clip(np.round(rng.normal(0, 0.05, (S, 320 * 125)) * 32767), -32768, 32767).astype(np.int16)
    for k in range(125):
        total += sess.write(base[:, 320 * k: 320 * (k + 1)]).shape[1]
    assert total == sess.windows == (320 * 125 - 160) // 256 + 1 - 62
