"""GPU parity: normalize_mfcc for every shape and method (extract_mfcc.py:47-88) and the explicit-context extract_mfcc."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mfcc as om

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(13, 40), (13, 63), (13, 1), (13, 2), (5, 13, 282), (13, 513), (3, 13, 4097)])
@pytest.mark.parametrize("method", ["cmvn", "standardization", "minmax", "other"])
def test_normalize_mfcc_any_shape_matches_the_reference_formula(cuda_device, shape, method):
    import ww_b200

    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 7.0 - 3.0
    if x.numel() and x.dim() >= 2:
        x[..., 0, :] = 2.5                               # a constant row: std == 0 -> 1 / max == min
    want = om.normalize_mfcc(x, method).numpy()          # the reference's formula on the CPU (oracle)
    got_dev = ww_b200.normalize_mfcc(x.to(cuda_device), method)
    got_cpu = ww_b200.normalize_mfcc(x, method)          # CPU tensor in -> CPU tensor out, computed on the GPU
    assert got_dev.is_cuda and not got_cpu.is_cuda and got_dev.shape == x.shape
    np.testing.assert_array_equal(got_dev.cpu().numpy(), got_cpu.numpy())
    np.testing.assert_allclose(got_cpu.numpy(), want, rtol=2e-5, atol=2e-5, equal_nan=True)


def test_normalize_rows_in_place_and_strided(cuda_device):
    from ww_b200 import _lib as L

    ctx = L.get_context(0)
    x = torch.randn(64, 1000, device=cuda_device)
    want = (x[:, :700] - x[:, :700].mean(1, keepdim=True)) / (x[:, :700].std(1, keepdim=True) + 1e-8)
    tail = x[:, 700:].clone()
    ctx.check(ctx.lib.ww_normalize_rows(ctx.h, L.ptr(x), 64, 700, 1000, L.NORM_STANDARD, L.ptr(x), L.cur_stream(cuda_device)),
              "ww_normalize_rows")
    torch.cuda.synchronize()
    assert torch.allclose(x[:, :700], want, atol=2e-5) and torch.equal(x[:, 700:], tail)
    assert ctx.lib.ww_normalize_rows(ctx.h, L.ptr(x), 64, 700, 600, L.NORM_STANDARD, L.ptr(x), None) == -1
    assert ctx.lib.ww_normalize_rows(ctx.h, L.ptr(x), 64, 700, 1000, 7, L.ptr(x), None) == -1


def test_extract_mfcc_with_an_explicit_context_equals_the_shim(cuda_device):
    from ww_b200 import _lib as L

    lib = L.load_library()
    rng = np.random.default_rng(0)
    sig = rng.normal(0, 0.1, 16000).astype(np.float32)
    ctx = L.Context(0)                                    # a context of its own, not the process-wide one
    try:
        a = lib.ww_extract_mfcc_ctx(ctx.h, sig.ctypes.data_as(C.c_void_p), 16000, 16000, 320, 256, 512, 40, 13)
        b = lib.ww_extract_mfcc(sig.ctypes.data_as(C.c_void_p), 16000, 16000, 320, 256, 512, 40, 13)
        assert a and b
        T = lib.ww_num_frames(L.FEAT_ESP, 16000)
        fa = np.ctypeslib.as_array(a, shape=(T * 13,)).copy()
        fb = np.ctypeslib.as_array(b, shape=(T * 13,)).copy()
        lib.ww_free_mfcc(a)
        lib.ww_free_mfcc(b)
        np.testing.assert_array_equal(fa, fb)
        assert not lib.ww_extract_mfcc_ctx(ctx.h, sig.ctypes.data_as(C.c_void_p), 100, 16000, 320, 256, 512, 40, 13)
        assert not lib.ww_extract_mfcc_ctx(None, sig.ctypes.data_as(C.c_void_p), 16000, 16000, 320, 256, 512, 40, 13)
    finally:
        ctx.close()
