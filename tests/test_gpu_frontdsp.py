"""GPU parity for the SURVEY.md 8f rank 3/4 rows: TDM down-mix (bit-exact), waveform augmentation (fp32, 1e-6),
WAV files -> pinned int16 batch -> scores."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn as ocnn
from oracle import frontdsp as ofd
from oracle import mfcc as omfcc
from oracle import wav as owav

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("B,n", [(1, 320), (3, 16000), (5, 1001), (2, 7), (1, 0)])
def test_tdm_downmix_bit_exact(cuda_device, B, n):
    """cpp:103-121: int16 wrap of the mix, floor shifts; vector path (aligned) and scalar tails."""
    import ww_b200

    rng = np.random.default_rng(100 + n)
    x = rng.integers(-32768, 32768, size=(B, 12 * n), dtype=np.int16)
    if n >= 3:
        x[:, :12] = 32767           # mix overflows int16 and wraps
        x[:, 12:24] = -32768
    got = ww_b200.tdm_downmix(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    want = ofd.tdm_downmix(x)
    assert got.shape == (B, n) and got.dtype == np.int16
    np.testing.assert_array_equal(got, want)
    if n:
        np.testing.assert_array_equal(got[0], ofd.tdm_downmix_c(x[0]))


def test_tdm_downmix_unaligned_rows_and_stream(cuda_device):
    """Row stride not a multiple of 8 elements (scalar path) and a 60 s stream processed in one call == 20 ms blocks."""
    import ww_b200

    rng = np.random.default_rng(9)
    base = torch.from_numpy(rng.integers(-32768, 32768, size=(4, 12 * 500 + 3), dtype=np.int16)).to(cuda_device)
    view = base[:, 1:1 + 12 * 500]          # misaligned start: the library must take the scalar path
    got = ww_b200.tdm_downmix(view.contiguous()).cpu().numpy()
    np.testing.assert_array_equal(got, ofd.tdm_downmix(view.cpu().numpy()))
    n = 16000 * 60
    s = rng.integers(-20000, 20000, size=12 * n, dtype=np.int16)
    whole = ww_b200.tdm_downmix(torch.from_numpy(s).to(cuda_device)).cpu().numpy()
    assert whole.shape == (n,)
    blocks = np.concatenate([ofd.tdm_downmix(s[i:i + 3840]) for i in range(0, 3840 * 50, 3840)])
    np.testing.assert_array_equal(whole[:blocks.size], blocks)
    np.testing.assert_array_equal(whole, ofd.tdm_downmix_c(s))


def test_tdm_then_frontend_matches_oracle_chain(cuda_device):
    """TDM capture -> mono 16 kHz -> MFCC: the chain record_task feeds, against the oracle chain."""
    import ww_b200

    rng = np.random.default_rng(3)
    t = np.arange(48000) / 48000.0
    mic = (8000 * np.sin(2 * np.pi * 700 * t) + 500 * rng.standard_normal(48000)).astype(np.int16)
    tdm = np.stack([mic, (mic // 3).astype(np.int16), np.roll(mic, 5), np.zeros_like(mic)], axis=1).reshape(-1)
    pcm = ww_b200.tdm_downmix(torch.from_numpy(tdm).to(cuda_device))
    feats = ww_b200.mfcc_batch(pcm[None]).cpu().numpy()
    want = omfcc.mfcc_torchaudio(omfcc.pcm16_to_float(ofd.tdm_downmix(tdm)[None])).numpy()
    assert np.abs(feats - want).max() < 1e-3


def test_augment_matches_reference_golden_and_oracle(cuda_device):
    """augment_audio_waveform: golden = the reference's own function output; tolerance 1e-6 abs (fp32 lerp order)."""
    import ww_b200

    d = np.load(os.path.join(ROOT, "tests", "golden", "wav_cases.npz"))
    got = ww_b200.augment_batch(torch.from_numpy(d["aug_in"]).to(cuda_device)).cpu().numpy()
    assert got.shape == (1, 5, 16000)
    assert np.abs(got - d["aug_out"]).max() <= 1e-6
    np.testing.assert_array_equal(got[:, 0], d["aug_in"])
    np.testing.assert_array_equal(got[:, 3:], d["aug_out"][:, 3:])   # volume variants are exact
    rng = np.random.default_rng(11)
    x = (rng.random((37, 16000), dtype=np.float32) * 2 - 1)
    got = ww_b200.augment_batch(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    assert np.abs(got - ofd.augment_waveform(x)).max() <= 1e-6


def test_wav_files_to_scores(cuda_device, xiaoa_sd, tmp_path):
    """WAV files (shipped-clip PCM from the golden set, with junk chunks) -> load_wav_batch -> score_host:
    same logits/decisions as scoring the PCM directly and as the oracle."""
    import ww_b200

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_features.npz"))
    pcm = g["pcm"]
    paths = []
    for i in range(pcm.shape[0]):
        n = 16000 if i % 3 else 9000 + 500 * i          # some shorter files (zero padded by the loader)
        body = owav.wav_bytes(pcm[i, :n])
        if i % 2:
            body = body[:36] + b"LIST" + (10).to_bytes(4, "little") + bytes(10) + body[36:]
        p = tmp_path / f"c{i}.wav"
        p.write_bytes(body)
        paths.append(str(p))
    batch, infos, st = ww_b200.load_wav_batch(paths)
    assert (st == 0).all() and batch.is_pinned()
    want_pcm = pcm.copy()
    for i in range(pcm.shape[0]):
        if i % 3 == 0:
            want_pcm[i, 9000 + 500 * i:] = 0
    np.testing.assert_array_equal(batch.numpy(), want_pcm)
    scorer = ww_b200.WakeWordScorer(xiaoa_sd, device=0)
    lg_h, dec_h = scorer.score_host(batch)
    lg_d, dec_d = scorer.score(torch.from_numpy(want_pcm).to(cuda_device))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(np.asarray(dec_h), dec_d.cpu().numpy())
    ref_f = omfcc.mfcc_torchaudio(omfcc.pcm16_to_float(want_pcm)).numpy()
    ref_l = ocnn.forward_torch(omfcc.normalize_mfcc(ref_f, "cmvn").numpy(), xiaoa_sd)
    assert np.abs(np.asarray(lg_h) - ref_l).max() < 1e-2
    margin = np.abs(ref_l[:, 0]) > 1e-3
    assert (np.asarray(dec_h).astype(bool) == ocnn.decide_python(ref_l[:, 0]))[margin].all()


def test_extract_features_dropin(cuda_device, tmp_path):
    """ww_b200.extract_features (the reference's signature, extract_mfcc.py:123) over a directory of WAV files:
    equals the goldens produced by the reference's OWN extract_features (tests/golden/make_golden.py); with
    augmentation the deterministic variants equal the oracle chain."""
    import ww_b200

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_features.npz"))
    n = 6
    for i in range(n):
        (tmp_path / f"clip_{i:02d}.wav").write_bytes(owav.wav_bytes(g["pcm"][i]))
    (tmp_path / "notes.txt").write_text("not a wav")
    feats, labels = ww_b200.extract_features(str(tmp_path), label=1, add_noise_to_pad=False, augment_audio=False)
    assert len(feats) == n and all(int(l) == 1 for l in labels)
    order = [int(name[5:7]) for name in os.listdir(tmp_path) if name.endswith(".wav")]
    got = torch.stack(feats).cpu().numpy()
    assert got.shape == (n, 13, 63)
    assert np.abs(got - g["mfcc_cmvn"][order]).max() < 2e-3
    feats5, _ = ww_b200.extract_features(str(tmp_path), add_noise_to_pad=False, augment_audio=True)
    assert len(feats5) == 5 * n
    got5 = torch.stack(feats5).cpu().numpy().reshape(n, 5, 13, 63)
    aug = ofd.augment_waveform(omfcc.pcm16_to_float(g["pcm"][order]))
    for v in (0, 2, 3, 4):   # variant 1 is re-padded with unseeded noise by the reference
        want = omfcc.normalize_mfcc(omfcc.mfcc_torchaudio(aug[:, v]).numpy(), "cmvn").numpy()
        assert np.abs(got5[:, v] - want).max() < 5e-3, v


def test_firmware_chain_tdm_sessions_int8(cuda_device, xiaoa_sd):
    """The firmware's whole live path as a batch service: read_mic's TDM chunks (20 ms = 3840 int16 per stream) ->
    mix + decimator -> MFCC -> int8 + device CMVN -> int8 model -> hits, through push/poll sessions.  Pushing TDM
    chunks must equal pushing the oracle's down-mixed PCM, and both must equal whole-stream scoring."""
    import ww_b200

    rng = np.random.default_rng(17)
    n_streams, seconds = 3, 3
    n16 = 16000 * seconds
    t = np.arange(3 * n16) / 48000.0
    tdm = np.zeros((n_streams, 3 * n16, 4), dtype=np.int16)
    for k in range(n_streams):
        mic = (6000 * np.sin(2 * np.pi * (300 + 170 * k) * t) + 900 * rng.standard_normal(3 * n16)).astype(np.int16)
        tdm[k, :, 0] = mic
        tdm[k, :, 1] = mic // 4
        tdm[k, :, 2] = np.roll(mic, 7)
        tdm[k, :, 3] = rng.integers(-100, 100, size=3 * n16)
    tdm = tdm.reshape(n_streams, -1)
    pcm = ofd.tdm_downmix(tdm)                                    # oracle mono 16 kHz [n_streams, n16]
    a = ww_b200.StreamSession(xiaoa_sd, n_streams, max_chunk_samples=320, cmvn="device", cnn_impl="int8")
    got_a = [a.write_tdm(tdm[:, 12 * i:12 * (i + 320)]) for i in range(0, n16, 320)]
    hits_a = a.poll()
    a.close()
    b = ww_b200.StreamSession(xiaoa_sd, n_streams, max_chunk_samples=320, cmvn="device", cnn_impl="int8")
    got_b = [b.write(pcm[:, i:i + 320]) for i in range(0, n16, 320)]
    hits_b = b.poll()
    b.close()
    la = np.concatenate([g for g in got_a if g.shape[1]], axis=1)
    lb = np.concatenate([g for g in got_b if g.shape[1]], axis=1)
    np.testing.assert_array_equal(la, lb)
    assert [tuple(h) for h in hits_a] == [tuple(h) for h in hits_b]
    ss = ww_b200.StreamScorer(xiaoa_sd, cmvn="device", cnn_impl="int8")
    for k in range(n_streams):
        _, lg = ss.score(torch.from_numpy(pcm[k]).to(cuda_device))
        np.testing.assert_array_equal(la[k], lg.cpu().numpy()[:la.shape[1]])


def test_score_wav_dir_float_and_device_paths(cuda_device, xiaoa_sd, tmp_path):
    """hello_world_main.cpp's test_model loop (WAV directory -> tally) on both decision paths."""
    import ww_b200

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_features.npz"))
    n = g["pcm"].shape[0]
    for i in range(n):
        (tmp_path / f"w{i:02d}.wav").write_bytes(owav.wav_bytes(g["pcm"][i]))
    names, logits, dec, pos = ww_b200.score_wav_dir(str(tmp_path), xiaoa_sd)
    assert names == [f"w{i:02d}.wav" for i in range(n)] and pos == int(dec.sum())
    assert np.abs(logits - g["logits"]).max() < 1e-2                      # the reference's own logits
    margin = np.abs(g["logits"][:, 0]) > 1e-2
    assert (dec.astype(bool) == (g["logits"][:, 0] > 0))[margin].all()
    names2, lq, dec2, pos2 = ww_b200.score_wav_dir(str(tmp_path), xiaoa_sd, device_path=True)
    out_q, dec_ref = ww_b200.score_clips_int8(xiaoa_sd, torch.from_numpy(g["pcm"]).to(cuda_device))
    np.testing.assert_array_equal(lq * 8.0, out_q.cpu().numpy().astype(np.float32))
    np.testing.assert_array_equal(dec2, dec_ref.cpu().numpy())


def test_files_to_decisions_pipeline_equals_load_then_score(cuda_device, xiaoa_sd, tmp_path, monkeypatch):
    """ww_score_wav_files (reader threads fill one pinned batch while the GPU scores the other) over several batches,
    odd batch count, a short file, a file with extra chunks and a missing file: the same logits / decisions as
    load_wav_batch followed by score_host (ml_models/src/extract_mfcc.py:151-176 walks a directory the same way)."""
    import ww_b200
    from ww_b200 import _lib as L
    from ww_b200.model import WakeWordScorer

    rng = np.random.default_rng(21)
    n = 333
    pcm = np.clip(np.round(rng.normal(0, 0.1, (n, 16000)) * 32767), -32768, 32767).astype(np.int16)
    paths = []
    for i in range(n):
        p = tmp_path / f"f{i:04d}.wav"
        ww_b200.write_wav(str(p), pcm[i, : (9000 if i % 50 == 7 else 16000)])
        if i % 40 == 3:   # a LIST chunk in front of "data"
            body = p.read_bytes()
            p.write_bytes(body[:36] + b"LIST" + (6).to_bytes(4, "little") + bytes(6) + body[36:])
        paths.append(str(p))
    monkeypatch.setenv("WW_HOST_CHUNK_CLIPS", "64")      # 6 batches of 64: the two staging buffers rotate three times
    ctx = L.Context(0)
    try:
        sc = WakeWordScorer(xiaoa_sd, device=0)
        sc.ctx = ctx                                      # a context of its own with the small batch size
        sc._key = ("pipeline-test", 0)
        logits, dec, infos, st, stats = ww_b200.score_wav_files(paths, sc, threads=4)
        assert (st == 0).all() and len(infos) == n and stats["total_s"] > 0
        batch, _, _ = ww_b200.load_wav_batch(paths)
        want_l, want_d = sc.score_host(batch)
        np.testing.assert_array_equal(logits, want_l)
        np.testing.assert_array_equal(dec, want_d)
        # a missing file: counted, scored as silence, the others unaffected
        bad = paths[:100] + [str(tmp_path / "nope.wav")] + paths[100:200]
        with pytest.raises(ww_b200.WWError):
            ww_b200.score_wav_files(bad, sc, threads=3)
        l2, d2, _, st2, _ = ww_b200.score_wav_files(bad, sc, threads=3, strict=False)
        assert st2[100] != 0 and (np.delete(st2, 100) == 0).all()
        np.testing.assert_array_equal(np.delete(l2, 100, axis=0), want_l[:200])
        silence = sc.score_host(np.zeros((1, 16000), np.int16))[0]
        np.testing.assert_array_equal(l2[100], silence[0])
    finally:
        ctx.close()
