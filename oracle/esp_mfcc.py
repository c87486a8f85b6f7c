"""ctypes access to the C-MFCC checkers -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  port : oracle/c/libesp_mfcc_port.so  (plain-C restatement of main/esp_mfcc/mfcc.c:431-527)
  ref  : oracle/_ref/libesp_mfcc_ref.so (the reference's own mfcc.c compiled against shims and
         esp-dsp stand-ins; exists only when it was built in the container that has /root/reference)
Both realise the intended FFT math; the FFT stage is "parity unpinned" (see esp_mfcc_port.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "c", "libesp_mfcc_port.so")
REF_SO = os.path.join(_HERE, "_ref", "libesp_mfcc_ref.so")


def build():
    """Compile the C checkers (gcc).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True, capture_output=True)


def n_frames(n_samples, frame=320, hop=256):
    return (n_samples - frame) // hop + 1


def esp_mfcc_port(signal):
    """signal: float32 [N] -> [num_frames, 13] via the plain-C restatement."""
    if not os.path.exists(PORT_SO):
        build()
    lib = C.CDLL(PORT_SO)
    lib.esp_mfcc_port.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    x = np.ascontiguousarray(signal, dtype=np.float32)
    T = n_frames(len(x))
    out = np.zeros((T, 13), dtype=np.float32)
    n = lib.esp_mfcc_port(x.ctypes.data, len(x), 16000, 320, 256, 512, 40, 13, out.ctypes.data)
    if n != T:
        raise RuntimeError(f"esp_mfcc_port returned {n}")
    return out


def have_ref():
    return os.path.exists(REF_SO)


def esp_mfcc_ref(signal):
    """signal: float32 [N] -> [num_frames, 13] via the reference's own extract_mfcc (oracle/_ref)."""
    lib = C.CDLL(REF_SO)
    lib.extract_mfcc.argtypes = [C.c_void_p] + [C.c_int] * 7
    lib.extract_mfcc.restype = C.POINTER(C.c_float)
    lib.free_mfcc.argtypes = [C.POINTER(C.c_float)]
    x = np.ascontiguousarray(signal, dtype=np.float32)
    p = lib.extract_mfcc(x.ctypes.data, len(x), 16000, 320, 256, 512, 40, 13)
    if not p:
        return None
    T = n_frames(len(x))
    out = np.ctypeslib.as_array(p, shape=(T * 13,)).copy().reshape(T, 13)
    lib.free_mfcc(p)
    return out
