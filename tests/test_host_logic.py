"""CPU: host-side pieces of the package (no GPU): ONNX reader, padding/augmentation glue, module layout."""
import os
import struct

import numpy as np
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


def test_normalize_mfcc_cpu_formula_matches_oracle():
    from oracle import mfcc as om

    x = torch.randn(13, 40)
    for method in ("cmvn", "standardization", "minmax", "other"):
        np.testing.assert_allclose(ww_b200.normalize_mfcc(x, method).numpy(), om.normalize_mfcc(x, method).numpy(),
                                   atol=1e-6)


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
