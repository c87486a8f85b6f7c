"""The C ABI consumed from plain C (gcc, no CUDA headers): builds tests/c/abi_smoke.c against libwwb200.so."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "esp32-wake-word_b200")
EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def _build():
    import __graft_entry__ as ge

    ge.build()
    subprocess.run(["gcc", "-O1", "-std=c11", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L" + PKG, "-lwwb200",
                    "-Wl,-rpath," + PKG], check=True)


def test_c_program_links_and_host_checks_pass():
    _build()
    import torch

    r = subprocess.run([EXE], capture_output=True, text=True)
    assert "frames(py,16000)=63 frames(esp,16000)=62" in r.stdout
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "ww_create rc=-2" in r.stdout  # WW_ERR_CUDA: no CPU fallback


@pytest.mark.gpu
def test_c_program_scores_clips_like_python(cuda_device, xiaoa_sd, tmp_path):
    import torch

    import ww_b200
    from oracle import esp_mfcc as oesp
    from oracle import mfcc as om

    _build()
    n = 300
    pcm = om.synth_clips_int16(n, seed=99)
    wfile, pfile, ofile = tmp_path / "w.bin", tmp_path / "p.bin", tmp_path / "o.bin"
    np.concatenate([xiaoa_sd[k].astype("<f4").ravel() for k in
                    ("conv_layers.0.weight", "conv_layers.3.weight", "conv_layers.6.weight", "classifier.0.weight",
                     "classifier.2.weight")]).tofile(wfile)
    pcm.astype("<i2").tofile(pfile)
    r = subprocess.run([EXE, "gpu", str(wfile), str(pfile), str(n), str(ofile)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(ofile, dtype=np.uint8)
    logits = raw[: 4 * n].view("<f4")
    dec = raw[4 * n: 5 * n]
    mf = raw[5 * n:].view("<f4").reshape(62, 13)
    sc = ww_b200.WakeWordScorer(xiaoa_sd, device=0, cnn_impl="tensor")
    lp, dp = sc.score(torch.from_numpy(pcm).to(cuda_device))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(logits, lp.cpu().numpy()[:, 0])
    np.testing.assert_array_equal(dec, dp.cpu().numpy())
    assert np.abs(mf - oesp.esp_mfcc_port(om.pcm16_to_float(pcm[:1])[0])).max() < 2e-3
