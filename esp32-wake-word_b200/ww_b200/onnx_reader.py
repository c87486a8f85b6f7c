"""Minimal ONNX initializer reader (no `onnx` package in this image).

Reads only what the wake-word path needs from an ONNX file: the graph's
initializer tensors (name, dims, fp32 payload).  The reference ships its CNN
as `ml_models/xiaoa.onnx` (export of `LightweightKWS(num_classes=1)`,
reference ml_models/src/wakeModel.py:4-34, used at ml_models/main.py:14).

Protobuf wire format, fields used:
  ModelProto.graph            = field 7 (len-delimited)
  GraphProto.initializer      = field 5 (len-delimited, repeated TensorProto)
  TensorProto.dims            = field 1 (varint, repeated, maybe packed)
  TensorProto.data_type       = field 2 (varint; 1 == FLOAT)
  TensorProto.float_data      = field 4 (packed fixed32)
  TensorProto.name            = field 8 (bytes)
  TensorProto.raw_data        = field 9 (bytes)
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, Tuple

import numpy as np


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims = []
    dtype = 0
    name = ""
    raw = None
    floats = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(val)
            else:  # packed
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif fno == 2:
            dtype = val
        elif fno == 4:
            if wt == 2:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", val))
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif fno == 8:
            name = val.decode("utf-8")
        elif fno == 9:
            raw = val
    if dtype != 1:
        return name, None  # only FLOAT initializers matter here
    if raw is not None:
        arr = np.frombuffer(raw, dtype="<f4").copy()
    else:
        arr = np.asarray(floats, dtype=np.float32)
    return name, arr.reshape(dims) if dims else arr


def read_initializers(path: str) -> Dict[str, np.ndarray]:
    """Return {initializer name: fp32 ndarray} for every FLOAT initializer."""
    with open(path, "rb") as f:
        buf = f.read()
    out: Dict[str, np.ndarray] = {}
    for fno, wt, val in _fields(buf):
        if fno == 7 and wt == 2:  # graph
            for gfno, gwt, gval in _fields(val):
                if gfno == 5 and gwt == 2:
                    name, arr = _tensor(gval)
                    if arr is not None:
                        out[name] = arr
    return out


def load_kws_state_dict(path: str) -> Dict[str, np.ndarray]:
    """Map an exported LightweightKWS ONNX file to torch `state_dict` keys.

    The export keeps conv weights as `conv_layers.{0,3,6}.weight` [O,I,3] and
    stores the two bias-free Linear layers as MatMul initialisers with the
    weight TRANSPOSED to [I,O]; this returns them in
    torch layout [O,I] under `classifier.{0,2}.weight`.
    """
    init = read_initializers(path)
    sd: Dict[str, np.ndarray] = {}
    for k in ("conv_layers.0.weight", "conv_layers.3.weight", "conv_layers.6.weight"):
        if k not in init:
            raise KeyError(f"{path}: initializer {k} not found (have {sorted(init)})")
        sd[k] = np.ascontiguousarray(init[k], dtype=np.float32)
    # the MatMul weights are the remaining 2-D initialisers, in graph order
    mats = [(k, v) for k, v in init.items() if v.ndim == 2 and k not in sd]
    fc1 = [v for _, v in mats if v.shape[0] == sd["conv_layers.6.weight"].shape[0]]
    if len(fc1) != 1:
        raise ValueError(f"{path}: cannot identify classifier.0 MatMul weight")
    fc1 = fc1[0]
    fc2 = [v for _, v in mats if v.shape[0] == fc1.shape[1] and v is not fc1]
    if len(fc2) != 1:
        raise ValueError(f"{path}: cannot identify classifier.2 MatMul weight")
    sd["classifier.0.weight"] = np.ascontiguousarray(fc1.T, dtype=np.float32)
    sd["classifier.2.weight"] = np.ascontiguousarray(fc2[0].T, dtype=np.float32)
    return sd
