#!/usr/bin/env python
"""Golden log lines of the reference's OWN analyze_mfcc_range (main/esp_mfcc/mfcc.c:530-553).

Runs in the build container only: oracle/_ref/libesp_mfcc_ref.so is the reference's mfcc.c compiled from
/root/reference (oracle/c/Makefile); the shim esp_log.h keeps the last ESP_LOGI / ESP_LOGE line in a buffer, which is
what this script records for a handful of crafted feature arrays.  Output: tests/golden/analyze_range.npz."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libesp_mfcc_ref.so"))
    lib.analyze_mfcc_range.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.analyze_mfcc_range.restype = None
    info = (C.c_char * 512).in_dll(lib, "shim_last_info")
    err = (C.c_char * 512).in_dll(lib, "shim_last_error")
    rng = np.random.default_rng(11)
    feats = np.load(os.path.join(HERE, "ref_features.npz"))
    key = [k for k in feats.files if feats[k].ndim >= 2 and feats[k].dtype == np.float32][0]
    cases = {
        "ref_features": feats[key].astype(np.float32).ravel()[: 13 * 63 * 4],
        "random_wide": (rng.standard_normal(5000) * 40 - 20).astype(np.float32),
        "with_nan_inf": np.array([1.5, np.nan, -3.25, np.inf, 7.0, -np.inf, 0.125], np.float32),
        "all_invalid": np.array([np.nan, np.inf, -np.inf], np.float32),
        "single": np.array([-87.377], np.float32),
        "large_sum": np.full(200000, 16.1, np.float32),     # the float accumulator loses bits here: order matters
    }
    out = {}
    for name, x in cases.items():
        info.value = b""
        err.value = b""
        lib.analyze_mfcc_range(x.ctypes.data_as(C.c_void_p), x.size, name.encode())
        line = info.value.decode() or ("E " + err.value.decode())
        out["x_" + name] = x
        out["line_" + name] = np.array(line)
        print(line)
    np.savez_compressed(os.path.join(HERE, "analyze_range.npz"), **out)


if __name__ == "__main__":
    main()
