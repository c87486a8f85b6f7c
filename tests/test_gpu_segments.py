"""GPU parity: a long stream split into halo segments (SURVEY.md 8e, configs[3]) stitches to the whole-stream result.

Reference semantics: esp_wake_word_detector.cpp:171-260 scores every 63-frame window of ONE endless stream; the
frame grid and the reflect padding belong to the stream, not to the piece of it a GPU happens to hold.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stream(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 0.03, n)
    for k in range(0, n // 16000, 3):
        s = k * 16000 + 5000
        x[s:s + 12000] += rng.normal(0, 0.2, min(12000, n - s))
    return np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("n_samples", [40 * 16000, 40 * 16000 + 1234, 63 * 256 + 300])
@pytest.mark.parametrize("cmvn,impl", [("python", "fp32"), ("device", "fp32"), ("python", "tensor"), ("device", "int8")])
@pytest.mark.parametrize("parts", [8, 3])
def test_segments_stitch_to_the_whole_stream(cuda_device, xiaoa_sd, n_samples, cmvn, impl, parts):
    import ww_b200
    from ww_b200 import shard

    pcm = _stream(n_samples, seed=n_samples % 97)
    x = torch.from_numpy(pcm).to(cuda_device)
    sc = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn=cmvn, cnn_impl=impl)
    feats, logits = sc.score(x)
    got_l, got_f = [], []
    for s0, s1, w0, nw in shard.stream_segments(n_samples, parts):
        if nw == 0:
            continue
        f, lg = sc.score_segment(x[s0:s1].clone(), s0, n_samples, w0, nw)   # .clone(): a buffer of its own, as on another GPU
        got_l.append(lg)
        got_f.append((w0, f))
    torch.cuda.synchronize()
    stitched = torch.cat(got_l)
    assert stitched.shape == logits.shape
    for w0, f in got_f:                                   # frames w0 .. w0 + nw + 61 of the whole stream's grid
        assert torch.equal(f, feats[:, w0:w0 + f.shape[1]])
    assert torch.equal(stitched, logits)                  # bit for bit, whatever the CNN implementation


def test_segment_without_its_halo_is_rejected(cuda_device, xiaoa_sd):
    import ww_b200

    n = 20 * 16000
    x = torch.from_numpy(_stream(n, 1)).to(cuda_device)
    sc = ww_b200.StreamScorer(xiaoa_sd, device=0)
    w0, nw = 100, 50
    s0, s1 = 256 * w0 - 161, 256 * (w0 + nw + 61) + 160
    sc.score_segment(x[s0:s1].clone(), s0, n, w0, nw)     # exactly the taps: fine (unaligned start -> non-TMA staging)
    with pytest.raises(ww_b200.WWError, match="left halo"):
        sc.score_segment(x[s0 + 1:s1].clone(), s0 + 1, n, w0, nw)
    with pytest.raises(ww_b200.WWError, match="right halo"):
        sc.score_segment(x[s0:s1 - 1].clone(), s0, n, w0, nw)
    with pytest.raises(ww_b200.WWError, match="left halo"):
        sc.score_segment(x[8:s1].clone(), 8, n, 0, nw)    # window 0 needs the stream's first samples (reflection)
    with pytest.raises(ww_b200.WWError):
        sc.score_segment(x[s0:s1].clone(), s0, n, w0, 0)  # no window


def test_unaligned_segment_equals_aligned_one(cuda_device, xiaoa_sd):
    import ww_b200

    n = 20 * 16000
    x = torch.from_numpy(_stream(n, 2)).to(cuda_device)
    sc = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn="python", cnn_impl="fp32")
    _, whole = sc.score(x)
    w0, nw = 333, 200
    s0, s1 = 256 * w0 - 161, 256 * (w0 + nw + 61) + 160
    for lead in (0, 1, 3, 7, 161 % 8):
        _, lg = sc.score_segment(x[s0 - lead:s1].clone(), s0 - lead, n, w0, nw)
        assert torch.equal(lg, whole[w0:w0 + nw])


@pytest.mark.parametrize("first", [8, 160, 168, 256, 320])
def test_session_first_chunk_shorter_than_a_frame(cuda_device, xiaoa_sd, first):
    """Frame 0 is reflect-padded on the left: its tap -160 is sample +160, so it needs 161 samples.  A first
    push of exactly 160 samples must not emit it early."""
    import ww_b200

    n = 63 * 256 + 4096
    pcm = _stream(n, 3)
    ref = ww_b200.StreamScorer(xiaoa_sd, device=0, cmvn="python", cnn_impl="fp32")
    _, want = ref.score(torch.from_numpy(pcm).to(cuda_device))
    sess = ww_b200.StreamSession(xiaoa_sd, 1, max_chunk_samples=4096, device=0, cmvn="python", cnn_impl="fp32")
    got = [sess.write(pcm[None, :first])]
    pos = first
    while pos + 8 <= n:
        step = min(4096, (n - pos) // 8 * 8)
        got.append(sess.write(pcm[None, pos:pos + step]))
        pos += step
    g = np.concatenate([a[0] for a in got], axis=0)
    assert g.shape[0] >= 1
    np.testing.assert_array_equal(g, want.cpu().numpy()[: g.shape[0]])
    sess.close()
