"""The one-kernel clip path (csrc/ww_fused.cuh: frontend pipelines and tcgen05 CNN groups on disjoint SMs of one
persistent launch, features handed over through an L2-resident ring) against the chunked launches and the oracle.

Both paths run the same arithmetic (mfcc_body / cnn_tc_body / cnn_fp32_kernel), so logits and decisions must be
IDENTICAL bit for bit -- for batch sizes that leave octets ragged, that wrap the 4096-clip ring several times, for
both CMVN styles, both PCM types, three classes, and when every window lands in the re-score list."""
import numpy as np
import pytest
import torch

from oracle import cnn as ocnn
from oracle import mfcc as om

pytestmark = pytest.mark.gpu


def _score(sc, pcm, fused):
    from ww_b200 import _lib as L

    sc._prep()
    sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED, fused), "ww_set_option")
    try:
        sc.ctx.lib.ww_tc_rescored_total(sc.ctx.h, 1)
        logits, dec = sc.score(pcm)
        torch.cuda.synchronize()
        resc = int(sc.ctx.lib.ww_tc_rescored_total(sc.ctx.h, 1))
    finally:
        sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED, 0), "ww_set_option")
    return logits, dec, resc


def _clips(n, seed, device):
    """n clips, 1536 distinct ones rolled by a clip-dependent offset so that no two rows are equal"""
    base = om.synth_clips_int16(min(n, 1536), seed=seed)
    if n <= base.shape[0]:
        return torch.from_numpy(base[:n]).to(device)
    t = torch.from_numpy(base).to(device)
    reps = (n + base.shape[0] - 1) // base.shape[0]
    out = torch.cat([torch.roll(t, shifts=37 * r, dims=1) for r in range(reps)], 0)[:n]
    return out.contiguous()


@pytest.mark.parametrize("cmvn,decision", [("python", "python"), ("device", "device")])
def test_fused_kernel_equals_chunked_launches(cuda_device, xiaoa_sd, cmvn, decision):
    import ww_b200

    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cmvn=cmvn, decision=decision, cnn_impl="tensor")
    for n in (1, 7, 8, 9, 63, 300, 4099, 20011):
        pcm = _clips(n, 100 + n, cuda_device)
        l0, d0, r0 = _score(sc, pcm, 0)
        l1, d1, r1 = _score(sc, pcm, 2)
        assert torch.equal(l0, l1), (n, (l0 - l1).abs().max().item())
        assert torch.equal(d0, d1), n
        assert r0 == r1, (n, r0, r1)
    # float PCM goes through the float instantiation
    pcm = _clips(5000, 7, cuda_device).to(torch.float32) / 32768.0
    l0, d0, _ = _score(sc, pcm, 0)
    l1, d1, _ = _score(sc, pcm, 2)
    assert torch.equal(l0, l1) and torch.equal(d0, d1)


def test_fused_auto_mode_and_oracle(cuda_device, xiaoa_sd):
    """WW_OPT_FUSED = 1 takes the one-kernel path from 2048 clips on; checked against the oracle."""
    import ww_b200

    n = 2560
    base = om.synth_clips_int16(n, seed=99)
    pcm = torch.from_numpy(base).to(cuda_device)
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    l0, d0, _ = _score(sc, pcm, 0)
    l1, d1, _ = _score(sc, pcm, 1)
    assert torch.equal(l0, l1) and torch.equal(d0, d1)
    feats_o = om.mfcc_torchaudio(om.pcm16_to_float(base)).numpy()
    logit_o = ocnn.forward_torch(om.normalize_mfcc(feats_o, "cmvn").numpy(), xiaoa_sd)[:, 0]
    assert np.abs(l1[:, 0].cpu().numpy() - logit_o).max() < 1e-2
    clear = np.abs(logit_o) > 1e-3
    assert (d1.cpu().numpy().astype(bool)[clear] == (logit_o > 0)[clear]).all()


def test_fused_every_window_rescored(cuda_device, xiaoa_sd):
    """Every clip is the same audio and the decision threshold is set to its exact logit: every window of every octet
    lands inside the guard band, is listed, copied out of the ring and re-scored by the exact kernel -- the worst case
    of the compact hand-over (and of the ring hold time)."""
    import ww_b200

    n = 9001
    one = torch.from_numpy(om.synth_clips_int16(1, seed=77)).to(cuda_device)
    pcm = one.repeat(n, 1).contiguous()
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    sc.threshold = float(ww_b200.WakeWordScorer(xiaoa_sd, device=0, cnn_impl="fp32").score(one)[0][0, 0].item())
    l0, d0, r0 = _score(sc, pcm, 0)
    l1, d1, r1 = _score(sc, pcm, 2)
    assert r0 == n and r1 == n, (r0, r1)
    assert torch.equal(l0, l1) and torch.equal(d0, d1)
    assert bool((l1 == l1[0]).all())


@pytest.mark.parametrize("chunk", [4096, 16384])
def test_l2_resident_chunks_equal_default_chunks(cuda_device, xiaoa_sd, chunk):
    """WW_OPT_L2_CHUNK_CLIPS: small frontend + CNN launch pairs through an L2-sized feature buffer, compact re-score list,
    one exact launch per 131 072 clips -- same bits as the 131 072-clip chunks, also when the list is long."""
    import ww_b200
    from ww_b200 import _lib as L

    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    for n, same in ((1, False), (chunk - 1, False), (3 * chunk + 77, False), (2 * chunk + 5, True)):
        if same:
            pcm = torch.from_numpy(om.synth_clips_int16(1, seed=78)).to(cuda_device).repeat(n, 1).contiguous()
            sc.threshold = float(ww_b200.WakeWordScorer(xiaoa_sd, device=0, cnn_impl="fp32").score(pcm[:1])[0][0, 0].item())
        else:
            pcm = _clips(n, 300 + n, cuda_device)
            sc.threshold = 0.0
        l0, d0, r0 = _score(sc, pcm, 0)
        sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_L2_CHUNK_CLIPS, chunk), "ww_set_option")
        try:
            l1, d1, r1 = _score(sc, pcm, 0)
        finally:
            sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_L2_CHUNK_CLIPS, 0), "ww_set_option")
        assert torch.equal(l0, l1) and torch.equal(d0, d1) and r0 == r1, (n, same, r0, r1)
        if same:
            assert r1 == n
    if chunk == 16384:
        # more clips than one exact re-score launch covers (131 072): the compact list is flushed and restarted mid-call
        n = 131072 + 16384 + 9
        pcm = _clips(n, 17, cuda_device)
        sc.threshold = 0.0
        l0, d0, r0 = _score(sc, pcm, 0)
        sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_L2_CHUNK_CLIPS, chunk), "ww_set_option")
        try:
            l1, d1, r1 = _score(sc, pcm, 0)
        finally:
            sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_L2_CHUNK_CLIPS, 0), "ww_set_option")
        assert torch.equal(l0, l1) and torch.equal(d0, d1) and r0 == r1


@pytest.mark.parametrize("cmvn,decision", [("python", "python"), ("device", "device")])
def test_deferred_rescore_of_the_default_handover_equals_per_chunk_rescore(cuda_device, xiaoa_sd, cmvn, decision):
    """WW_OPT_RESCORE_WINDOW_CLIPS: a call that spans several 131 072-clip chunks copies the windows inside the guard band
    to a compact buffer and re-scores them with ONE exact launch (per window of clips) -- same bits and the same re-score
    count as one exact launch per chunk, for a window that covers the call, a window of one chunk (flush between the
    chunks), and when EVERY window is listed (the compact buffer's worst case)."""
    import ww_b200
    from ww_b200 import _lib as L

    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cmvn=cmvn, decision=decision)
    n = 2 * 131072 + 4099
    for same in (False, True):
        if same:
            if cmvn != "python":
                continue
            pcm = torch.from_numpy(om.synth_clips_int16(1, seed=78)).to(cuda_device).repeat(n, 1).contiguous()
            sc.threshold = float(ww_b200.WakeWordScorer(xiaoa_sd, device=0, cnn_impl="fp32").score(pcm[:1])[0][0, 0].item())
        else:
            pcm = _clips(n, 23, cuda_device)
        out = {}
        try:
            for window in (0, 1 << 20, 1024):
                sc._prep()
                sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_RESCORE_WINDOW_CLIPS, window), "ww_set_option")
                out[window] = _score(sc, pcm, 0)
        finally:
            sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_RESCORE_WINDOW_CLIPS, 1 << 20), "ww_set_option")
        l0, d0, r0 = out[0]
        for window in (1 << 20, 1024):
            l1, d1, r1 = out[window]
            assert torch.equal(l0, l1) and torch.equal(d0, d1) and r0 == r1, (cmvn, same, window, r0, r1)
        if same:
            assert r0 == n
        else:
            assert 0 < r0 < n // 20


def test_fused_three_classes_and_cnn_sm_count(cuda_device):
    """random 3-class weights (the CTC keyword shape) and a non-default split of the SMs between the two roles"""
    import ww_b200
    from ww_b200 import _lib as L

    rng = np.random.default_rng(5)
    sd = {
        "conv_layers.0.weight": (rng.standard_normal((32, 13, 3)) * 0.2).astype(np.float32),
        "conv_layers.3.weight": (rng.standard_normal((64, 32, 3)) * 0.1).astype(np.float32),
        "conv_layers.6.weight": (rng.standard_normal((128, 64, 3)) * 0.08).astype(np.float32),
        "classifier.0.weight": (rng.standard_normal((64, 128)) * 0.1).astype(np.float32),
        "classifier.2.weight": (rng.standard_normal((3, 64)) * 0.2).astype(np.float32),
    }
    sc = ww_b200.WakeWordScorer(sd, device=0)
    pcm = _clips(6000, 11, cuda_device)
    l0, d0, r0 = _score(sc, pcm, 0)
    for sms in (3, 24):
        sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED_CNN_SMS, sms), "ww_set_option")
        try:
            l1, d1, r1 = _score(sc, pcm, 2)
        finally:
            sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED_CNN_SMS, 0), "ww_set_option")
        assert l1.shape == (6000, 3)
        assert torch.equal(l0, l1) and torch.equal(d0, d1) and r0 == r1


def test_fused_host_path(cuda_device, xiaoa_sd):
    """ww_score_clips_host (16 384-clip chunks, H2D / compute / D2H overlapped) over the one-kernel path"""
    import ww_b200
    from ww_b200 import _lib as L

    n = 40000
    pcm = _clips(n, 21, cuda_device).cpu().numpy()
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    l0, d0 = sc.score_host(pcm)
    sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED, 2), "ww_set_option")
    try:
        l1, d1 = sc.score_host(pcm)
    finally:
        sc.ctx.check(sc.ctx.lib.ww_set_option(sc.ctx.h, L.OPT_FUSED, 0), "ww_set_option")
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(d0, d1)
