#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference itself.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed from the reference (read-only, never copied into this repo):
  * ml_models/src/extract_mfcc.py  -- imported; `extract_features`, `normalize_mfcc`, `pad_audio` are
    called as they are.  `torchaudio.load` cannot decode here (torchcodec is absent), so it is replaced
    by a `wave`-based loader with the same return convention (float32 [1, N] = int16 / 32768, sr).
  * ml_models/src/wakeModel.py     -- imported; `LightweightKWS(num_classes=1)` with xiaoa.onnx weights.
  * ml_models/test.py:201-217 (`ctc_greedy_decode`) and ml_models/ctc.py:453-471
    (`decode_predictions`) -- the two modules train at import time and need librosa/matplotlib, so the
    two function bodies are lifted with `ast` and executed stand-alone.
Data parsed (not executed): ml_models/xiaoa.onnx (weights), ml_models/xiaoa.info (the shipped
known-answer vector), main/hello_world_main.cpp:50-132 (int8 MFCC dumps data1/data2).

Outputs (tests/golden/*.npz) are small and committed together with this script.
"""
import ast
import os
import re
import shutil
import sys
import tempfile
import wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))

from oracle import mfcc as omfcc  # noqa: E402
from ww_b200.onnx_reader import load_kws_state_dict  # noqa: E402


def wave_load(path):
    with wave.open(path, "rb") as w:
        n, ch, sr = w.getnframes(), w.getnchannels(), w.getframerate()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch).T
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0), sr


def write_wav(path, pcm16):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(np.asarray(pcm16, dtype="<i2").tobytes())


def import_reference():
    sys.path.insert(0, os.path.join(REF, "ml_models"))
    import torchaudio

    torchaudio.load = wave_load  # same return convention as torchaudio.load for 16-bit PCM
    import src.extract_mfcc as ref_mfcc
    import src.wakeModel as ref_model

    return ref_mfcc, ref_model


def lift_function(path, name):
    """Return the source of function `name` found anywhere in `path` (module is NOT imported)."""
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            return compile(ast.fix_missing_locations(mod), path, "exec")
    raise KeyError(name)


def parse_c_int_array(text, name):
    m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    return np.array([int(v) for v in re.findall(r"-?\d+", m.group(1))], dtype=np.int8)


def main():
    ref_mfcc, ref_model = import_reference()
    sd = load_kws_state_dict(os.path.join(REF, "ml_models", "xiaoa.onnx"))
    np.savez(os.path.join(HERE, "xiaoa_weights.npz"), **sd)

    model = ref_model.LightweightKWS(num_classes=1)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model.eval()

    # ---- features + logits through the reference's extract_features -------------------------------
    wav_dir = os.path.join(REF, "audio_data", "flash")
    real = sorted(f for f in os.listdir(wav_dir) if f.endswith(".wav"))[:6]
    synth = omfcc.synth_clips_int16(8, seed=1234)
    tmp = tempfile.mkdtemp()
    try:
        names, pcm = [], []
        for i, f in enumerate(real):
            a, sr = wave_load(os.path.join(wav_dir, f))
            assert sr == 16000
            p = np.round(a[0].numpy() * 32768.0).astype(np.int16)
            p = omfcc.pad_audio(p)  # zero pad / truncate to 16000 (add_noise_to_pad=False below)
            names.append("r%02d.wav" % i)
            pcm.append(p)
        for i in range(len(synth)):
            names.append("s%02d.wav" % i)
            pcm.append(synth[i])
        feats_raw, feats_cmvn, logits = [], [], []
        for n, p in zip(names, pcm):
            d = os.path.join(tmp, n[:-4])
            os.makedirs(d)
            write_wav(os.path.join(d, n), p)
            f_raw, _ = ref_mfcc.extract_features(d, label=1, add_noise_to_pad=False, augment_audio=False,
                                                 normalize_method="none")
            f_cm, _ = ref_mfcc.extract_features(d, label=1, add_noise_to_pad=False, augment_audio=False,
                                                normalize_method="cmvn")
            assert len(f_raw) == 1 and f_raw[0].shape == (13, 63)
            feats_raw.append(f_raw[0].numpy())
            feats_cmvn.append(f_cm[0].numpy())
            with torch.no_grad():
                logits.append(model(f_cm[0][None]).numpy()[0])
    finally:
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, "ref_features.npz"), pcm=np.stack(pcm), names=np.array(names),
                        mfcc=np.stack(feats_raw), mfcc_cmvn=np.stack(feats_cmvn), logits=np.stack(logits))

    # ---- the shipped known-answer vector (xiaoa.info) ----------------------------------------------
    info = open(os.path.join(REF, "ml_models", "xiaoa.info")).read()
    seg = info[info.index("test inputs value"):]
    vin = re.search(r"value: array\(\[(.*?)\]", seg, re.S).group(1)
    vin = np.array([int(v) for v in re.findall(r"-?\d+", vin)], dtype=np.int8)
    seg = info[info.index("test outputs value"):]
    vout = re.search(r"value: array\(\[(.*?)\]", seg, re.S).group(1)
    vout = np.array([int(v) for v in re.findall(r"-?\d+", vout)], dtype=np.int8)
    kat_in = vin[:63 * 13].reshape(63, 13)  # [frame][coef], exponent -4
    with torch.no_grad():
        kat_fp32 = model(torch.from_numpy(kat_in.T.astype(np.float32) / 16.0)[None]).numpy()[0]
    np.savez(os.path.join(HERE, "kat_xiaoa_info.npz"), input_q=kat_in, input_exponent=-4,
             output_q=vout[:1], output_exponent=-3, fp32_logit=kat_fp32)

    # ---- device MFCC dumps (hello_world_main.cpp data1 / data2) ------------------------------------
    cpp = open(os.path.join(REF, "main", "hello_world_main.cpp"), encoding="utf-8", errors="ignore").read()
    data2 = parse_c_int_array(cpp, "data2").reshape(13, 63)   # coef-major
    data1 = parse_c_int_array(cpp, "data1").reshape(63, 13)   # frame-major
    dumps = np.stack([data1.T, data2]).astype(np.float32)     # [2, 13, 63]
    z, q = omfcc.cmvn_device(dumps)
    with torch.no_grad():
        dl = model(torch.from_numpy(z)).numpy()
    np.savez(os.path.join(HERE, "device_dumps.npz"), mfcc_i8=dumps.astype(np.int8), cmvn_q=q, logits=dl)

    # ---- CTC decoders lifted from the reference scripts ----------------------------------------------
    ns = {"torch": torch, "print": lambda *a, **k: None}
    exec(lift_function(os.path.join(REF, "ml_models", "test.py"), "ctc_greedy_decode"), ns)
    from typing import List
    ns["List"] = List
    exec(lift_function(os.path.join(REF, "ml_models", "ctc.py"), "decode_predictions"), ns)
    rng = np.random.default_rng(2024)
    cases = []
    for (B, T, Cn) in ((16, 63, 3), (8, 63, 5), (4, 200, 30)):
        x = rng.normal(size=(B, T, Cn)).astype(np.float32)
        # make runs and blanks likely
        x[..., 0] += 0.8
        x = np.repeat(x[:, ::2], 2, axis=1)[:, :T]
        lp = torch.log_softmax(torch.from_numpy(x), dim=-1)
        chars = {i: chr(ord("a") + i - 1) if i else "_" for i in range(Cn)}

        class Dummy:
            idx_to_char = chars

        keep = [ns["ctc_greedy_decode"](Dummy(), lp[b], chars) for b in range(B)]
        coll = ns["decode_predictions"](Dummy(), lp)
        cases.append((lp.numpy(), keep, coll))
    np.savez_compressed(
        os.path.join(HERE, "ctc_decode.npz"),
        **{f"lp{i}": c[0] for i, c in enumerate(cases)},
        **{f"keep{i}": np.array(c[1]) for i, c in enumerate(cases)},
        **{f"collapse{i}": np.array(c[2]) for i, c in enumerate(cases)})
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
