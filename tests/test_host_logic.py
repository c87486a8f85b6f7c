"""CPU: host-side pieces of the package (no GPU): ONNX reader, padding/augmentation glue, module layout."""
import os
import struct

import numpy as np
import pytest
import torch

import ww_b200
from ww_b200 import onnx_reader


def _varint(n):
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _field(no, wt, payload):
    key = _varint((no << 3) | wt)
    return key + (_varint(len(payload)) + payload if wt == 2 else payload)


def _tensor(name, arr, raw=True):
    msg = b"".join(_field(1, 0, _varint(d)) for d in arr.shape)
    msg += _field(2, 0, _varint(1))
    msg += _field(8, 2, name.encode())
    if raw:
        msg += _field(9, 2, arr.astype("<f4").tobytes())
    else:
        msg += _field(4, 2, struct.pack(f"<{arr.size}f", *arr.ravel()))
    return msg


def test_onnx_reader_roundtrip(tmp_path, xiaoa_sd):
    tensors = {
        "conv_layers.0.weight": xiaoa_sd["conv_layers.0.weight"],
        "conv_layers.3.weight": xiaoa_sd["conv_layers.3.weight"],
        "conv_layers.6.weight": xiaoa_sd["conv_layers.6.weight"],
        "onnx::MatMul_23": xiaoa_sd["classifier.0.weight"].T.copy(),
        "onnx::MatMul_24": xiaoa_sd["classifier.2.weight"].T.copy(),
    }
    graph = b"".join(_field(5, 2, _tensor(k, v, raw=(i % 2 == 0))) for i, (k, v) in enumerate(tensors.items()))
    model = _field(1, 0, _varint(8)) + _field(7, 2, graph)
    p = tmp_path / "m.onnx"
    p.write_bytes(model)
    init = onnx_reader.read_initializers(str(p))
    assert set(init) == set(tensors)
    sd = onnx_reader.load_kws_state_dict(str(p))
    for k in xiaoa_sd:
        np.testing.assert_array_equal(sd[k], xiaoa_sd[k])


def test_module_structure_loads_reference_state_dict(xiaoa_sd):
    m = ww_b200.LightweightKWS(num_classes=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in xiaoa_sd.items()})
    assert sorted(m.state_dict()) == sorted(xiaoa_sd)
    assert sum(p.numel() for p in m.parameters()) == 40224
    assert ww_b200.LightweightKWS().classifier[2].out_features == 3  # reference default num_classes=3


def test_pad_audio_and_augment():
    a = torch.ones(1, 100)
    assert ww_b200.pad_audio(a, 160, add_noise_to_pad=False).shape == (1, 160)
    assert ww_b200.pad_audio(a, 160, add_noise_to_pad=False)[0, 100:].abs().sum() == 0
    assert ww_b200.pad_audio(a, 160, add_noise_to_pad=True)[0, 100:].abs().sum() > 0
    assert ww_b200.pad_audio(a, 50).shape == (1, 50)
    v = ww_b200.augment_audio_waveform(torch.rand(1, 16000) - 0.5)
    assert len(v) == 5 and all(x.shape == (1, 16000) for x in v)
    n = ww_b200.add_random_noise(torch.rand(1, 1000) - 0.5)
    assert n.abs().max() <= 1.0


def test_normalize_mfcc_has_no_cpu_path():
    """normalize_mfcc runs in libwwb200.so for every shape and method (tests/test_gpu_norm.py checks the numbers);
    without a GPU it raises instead of evaluating the formula with torch ops.  Unknown methods return the input,
    as the reference does (extract_mfcc.py:85-86)."""
    x = torch.randn(13, 40)
    assert ww_b200.normalize_mfcc(x, "other") is x
    if not torch.cuda.is_available():
        for method in ("cmvn", "standardization", "minmax"):
            with pytest.raises(ww_b200.WWError):
                ww_b200.normalize_mfcc(x, method)


def test_keyword_detector_accepts_the_reference_constructor():
    """ml_models/test.py:158-166: CTCKeywordDetector(model, char_to_idx, keywords, threshold=0.8); the model-less form
    scores log-probabilities the caller already has."""
    c2i = {"_": 0, "x": 1, "a": 2}
    model = torch.nn.Identity()
    a = ww_b200.CTCKeywordDetector(model, c2i, ["xa"])
    b = ww_b200.CTCKeywordDetector(c2i, ["xa"], 0.5)
    c = ww_b200.CTCKeywordDetector(model, c2i, ["xa"], threshold=0.7)
    assert a.model is model and a.threshold == 0.8 and a.idx_to_char == {0: "_", 1: "x", 2: "a"}
    assert b.model is None and b.threshold == 0.5 and c.threshold == 0.7
    assert a.calculate_confidence("xxa", "xa") == 0.9 and a.calculate_confidence("x", "xa") == 0.0
    with pytest.raises(TypeError):
        ww_b200.CTCKeywordDetector(c2i)
    with pytest.raises(ValueError):
        b.detect_keywords([])


def test_load_wav(tmp_path):
    import wave

    pcm = (np.arange(1000) - 500).astype("<i2")
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    a, sr = ww_b200.load_wav(str(p))
    assert sr == 16000 and a.shape == (1, 1000)
    np.testing.assert_array_equal(a[0].numpy(), pcm.astype(np.float32) / 32768.0)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "esp32-wake-word_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_generated_frontend_code_is_in_sync_with_its_generator(tmp_path):
    """csrc/ww_tables.h and ww_mel_py.inc are generated from the installed torchaudio (tools/gen_tables.py): regenerating
    into a scratch directory must reproduce the committed files byte for byte (modulo the torchaudio version line), and
    the work plan must cover every mel filter and DCT coefficient exactly once."""
    import subprocess
    import sys

    pytest.importorskip("torchaudio")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WW_GEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_tables.py")], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    def body(path):
        return [ln for ln in open(path).read().splitlines() if "GENERATED by" not in ln]

    csrc = os.path.join(root, "esp32-wake-word_b200", "csrc")
    for name in ("ww_tables.h", "ww_mel_py.inc"):
        assert body(os.path.join(tmp_path, name)) == body(os.path.join(csrc, name)), name + " is stale: run tools/gen_tables.py"

    sys.path.insert(0, os.path.join(root, "tools"))
    try:
        import gen_tables as gt
    finally:
        sys.path.pop(0)
    import re

    txt = open(os.path.join(csrc, "ww_tables.h")).read()
    start = [int(v) for v in re.search(r"WW_PY_MEL_START\[40\] = \{(.*?)\}", txt, re.S).group(1).replace("\n", " ").split(",") if v.strip()]
    length = [int(v) for v in re.search(r"WW_PY_MEL_LEN\[40\] = \{(.*?)\}", txt, re.S).group(1).replace("\n", " ").split(",") if v.strip()]
    ranges, warps = gt.plan(start, length)
    assert ranges[0][0] == 0 and ranges[-1][1] == 40 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert all(a % 2 == 0 for a, _ in ranges) and len(ranges) == 8
    assert sorted(q for qs in warps.values() for q in qs) == list(range(13))
    assert all(len({q % 2 for q in qs}) == 1 for qs in warps.values())


def test_extract_features_host_batch_keeps_channel_0_and_counts_real_samples(tmp_path):
    """Host part of extract_features (no GPU): mono, short and stereo files in one batch; a stereo file contributes
    its first channel (torchaudio.load -> [C, N], the reference keeps [0], extract_mfcc.py:154-172)."""
    import wave

    from ww_b200 import features

    rng = np.random.default_rng(8)
    mono = rng.integers(-20000, 20000, 16000).astype(np.int16)
    short = rng.integers(-20000, 20000, 5000).astype(np.int16)
    stereo = rng.integers(-20000, 20000, (12000, 2)).astype(np.int16)
    paths = []
    for name, data, ch in (("a", mono, 1), ("b", short, 1), ("c", stereo, 2)):
        p = str(tmp_path / (name + ".wav"))
        with wave.open(p, "wb") as w:
            w.setnchannels(ch)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(data.tobytes())
        paths.append(p)
    pcm, n_valid = features._load_clip_batch(paths)
    assert tuple(pcm.shape) == (3, 16000) and n_valid == [16000, 5000, 12000]
    np.testing.assert_array_equal(pcm[0].numpy(), mono)
    np.testing.assert_array_equal(pcm[1, :5000].numpy(), short)
    np.testing.assert_array_equal(pcm[2, :12000].numpy(), stereo[:, 0])
    assert int(pcm[1, 5000:].abs().sum()) == 0 and int(pcm[2, 12000:].abs().sum()) == 0
