"""CPU: the C-ABI library loads, exports every symbol include/ww_b200.h declares, and its host-only
entry points behave (no compute calls are made without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import stream as ostream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    import ww_b200

    return ww_b200.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ww_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ww_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 19
    from ww_b200._lib import EXPORTS

    assert names == set(EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_python_constants_follow_the_header():
    """the option ids and flag bits ww_b200 passes through ctypes are the ones include/ww_b200.h defines"""
    hdr = open(os.path.join(ROOT, "include", "ww_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    opts = {k: int(v) for k, v in re.findall(r"\bWW_(OPT_[A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    assert len(opts) >= 7
    from ww_b200 import _lib as L
    from ww_b200 import ctc as wctc

    for name, value in opts.items():
        assert getattr(L, name) == value, name
    assert int(re.search(r"#define\s+WW_CTC_BETA_IN_FWD\s+(\d+)", hdr).group(1)) == wctc._FLAG_BETA_IN_FWD


def test_num_frames(lib):
    assert lib.ww_num_frames(0, 16000) == 63       # torch.stft center=True
    assert lib.ww_num_frames(1, 16000) == 62       # mfcc.c:448
    assert lib.ww_num_frames(0, 57600000) == 225001
    assert lib.ww_num_frames(0, 256) == 0 and lib.ww_num_frames(1, 319) == 0
    assert lib.ww_num_frames(1, 320) == 1


def test_create_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.ww_create(C.byref(h), 0) == -2 and not h
    import ww_b200

    with pytest.raises(ww_b200.WWError):
        ww_b200.mfcc_batch(torch.zeros(1, 16000))
    with pytest.raises(ww_b200.WWError):
        ww_b200.LightweightKWS(1)(torch.zeros(1, 13, 63))


def test_workspace_bytes(lib):
    assert lib.ww_ctc_loss_workspace_bytes(63, 8, 2) == 2 * 8 * 63 * 5 * 4 + 16 + 8 * 16


def test_extract_mfcc_shim_rejects_bad_args(lib):
    x = np.zeros(100, np.float32)
    assert not lib.ww_extract_mfcc(x.ctypes.data, 100, 16000, 320, 256, 512, 40, 13)   # signal_len < frame
    assert not lib.ww_extract_mfcc(None, 16000, 16000, 320, 256, 512, 40, 13)         # NULL signal


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_stream_events_match_oracle(lib, seed):
    import ww_b200

    rng = np.random.default_rng(seed)
    n = 5000
    lg = rng.normal(-2.0, 1.5, size=(n, 1)).astype(np.float32)
    lg[rng.integers(0, n, 12)] += 6.0
    for refr in (0, 17, 313):
        want = ostream.events(lg, refractory=refr)
        got = ww_b200.events(lg, refractory=refr)
        assert got == want
    assert ww_b200.events(lg[:0]) == []
    # the very first window (frames 0..62) is never scored: esp_wake_word_detector.cpp:38-44
    lg2 = np.full((100, 1), -5.0, np.float32)
    lg2[0] = 9.0
    assert ww_b200.events(lg2) == []
    lg2[1] = 9.0
    assert ww_b200.events(lg2) == [1]


def test_refractory_frames():
    import ww_b200

    assert ww_b200.refractory_frames() == ostream.refractory_frames() == 313


def test_analyze_mfcc_range_logs_what_the_reference_logs(lib, capfd):
    """ww_analyze_mfcc_range (host only, no GPU) against the golden log lines of the reference's own
    analyze_mfcc_range (main/esp_mfcc/mfcc.c:530-553): same float accumulator, same text, NaN / Inf skipped."""
    import ctypes as C

    from oracle import mfcc as omfcc
    from ww_b200 import _lib as L

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "analyze_range.npz"))
    for name in [k[2:] for k in g.files if k.startswith("x_")]:
        x = np.ascontiguousarray(g["x_" + name], np.float32)
        out = L.MfccRange()
        n = lib.ww_analyze_mfcc_range(x.ctypes.data_as(C.c_void_p), x.size, name.encode(), C.byref(out))
        want = str(g["line_" + name])
        logged = capfd.readouterr().err.strip()
        if want.startswith("E "):
            assert n == 0 and out.valid == 0 and logged == want[2:]
        else:
            assert logged == want
            r = {"min": out.min_val, "max": out.max_val, "avg": out.avg, "valid": out.valid, "size": out.size}
            assert omfcc.analyze_range_line(name, r) == want and n == out.valid
    assert lib.ww_analyze_mfcc_range(None, 5, None, None) < 0          # the reference returns silently
    x = np.zeros(4, np.float32)
    assert lib.ww_analyze_mfcc_range(x.ctypes.data_as(C.c_void_p), 0, None, None) < 0
    import ww_b200
    assert ww_b200.analyze_mfcc_range(np.array([1.0, np.nan, 3.0], np.float32)) == \
        {"min": 1.0, "max": 3.0, "avg": 2.0, "valid": 2, "size": 3}


def test_ring_buffer_replays_the_references_own_test_scenario(lib):
    """ring_buffer_test_simple (main/ring_buffer/ring_buffer.c:120-200): create 10, write 3, read 3, write 7 (now
    full), read 3 -- with the intended keep-last-N / non-consuming-read semantics."""
    import ww_b200

    r = ww_b200.RingBuffer(10)
    r.write([1.1, 2.2, 3.3])
    assert len(r) == 3
    np.testing.assert_array_equal(r.read(3), np.array([1.1, 2.2, 3.3], np.float32))
    assert len(r) == 3                                           # read_rinbuffer takes a const ring: nothing consumed
    r.write([4.4, 5.5, 6.6, 7.7, 8.8, 9.9, 10.10])
    assert len(r) == 10
    np.testing.assert_array_equal(r.read(3), np.array([1.1, 2.2, 3.3], np.float32))
    r.write([11.0])                                              # full: the oldest value goes
    np.testing.assert_array_equal(r.read(10), np.array([2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9, 10.10, 11.0], np.float32))
    with pytest.raises(ww_b200.WWError):
        r.read(11)                                               # more than held: RINBUF_ERROR
    with pytest.raises(ww_b200.WWError):
        r.write([])
    r.close()
    with pytest.raises(ww_b200.WWError):
        ww_b200.RingBuffer(0)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ring_buffer_matches_the_oracle_model_on_random_traffic(lib, seed):
    import ww_b200

    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 200))
    r, m = ww_b200.RingBuffer(n), ostream.RingModel(n)
    for _ in range(300):
        k = int(rng.integers(1, 3 * n))                          # also writes longer than the ring (ring_buffer.c:63-66)
        x = rng.standard_normal(k).astype(np.float32)
        r.write(x)
        m.write(x)
        assert len(r) == m.count()
        q = int(rng.integers(1, n + 2))
        want = m.read(q)
        if want is None:
            with pytest.raises(ww_b200.WWError):
                r.read(q)
        else:
            np.testing.assert_array_equal(r.read(q), want)
    r.close()


def test_null_handles_are_rejected_without_touching_the_gpu(lib):
    """Every entry point that takes a context / session / ring checks it before anything else (no GPU here)."""
    import ctypes as C

    out = C.c_void_p()
    assert lib.ww_session_open(None, 4, 320, 1, 1, C.c_float(0.0), 64, 313, C.byref(out)) < 0 and not out.value
    x = np.zeros(320, np.int16)
    assert lib.ww_session_write(None, x.ctypes.data_as(C.c_void_p), 320) < 0
    assert lib.ww_session_write_tdm(None, x.ctypes.data_as(C.c_void_p), 24) < 0
    assert lib.ww_session_poll(None, None, 0) < 0
    assert lib.ww_session_windows(None) < 0
    lib.ww_session_close(None)                                   # like free(NULL)
    assert lib.ww_ring_write(None, x.ctypes.data_as(C.c_void_p), 4) < 0
    assert lib.ww_ring_read(None, x.ctypes.data_as(C.c_void_p), 4) < 0
    assert lib.ww_ring_count(None) < 0
    lib.ww_ring_delete(None)
    assert lib.ww_score_clips_host(None, x.ctypes.data_as(C.c_void_p), 0, 1, 1, 1, C.c_float(0.0), 1, None, None) < 0
