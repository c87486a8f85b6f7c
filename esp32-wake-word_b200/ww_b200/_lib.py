"""ctypes binding of libwwb200.so (include/ww_b200.h).

PyTorch is used for device memory and streams only: every call hands raw device
pointers and the current CUDA stream to the C ABI.  There is no CPU fallback:
if the library is missing, or no B200 is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# WW_B200_LIB selects another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("WW_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libwwb200.so")

WW_OK = 0
FEAT_PY, FEAT_ESP = 0, 1
PCM_S16, PCM_F32 = 0, 1
LAYOUT_COEF_MAJOR, LAYOUT_FRAME_MAJOR = 0, 1
CMVN_NONE, CMVN_PY, CMVN_DEVICE = 0, 1, 2
DECIDE_NONE, DECIDE_LOGIT, DECIDE_DEVICE = 0, 1, 2
DECODE_KEEP_REPEATS, DECODE_COLLAPSE = 0, 1
CNN_FP32, CNN_TENSOR, CNN_INT8 = 0, 1, 2
OPT_I8_IMPL = 1
OPT_GENERIC_FRONTEND = 2
OPT_FUSED = 3            # 0 chunked launches (default), 1 one persistent kernel from 2048 clips on, 2 always one kernel
OPT_FUSED_CNN_SMS = 4    # SMs given to the CNN role of the one-kernel path (0 = default)
OPT_L2_CHUNK_CLIPS = 5    # chunked tensor path with the features kept in L2: clips per launch pair (0 = off)
OPT_RESCORE_WINDOW_CLIPS = 6  # default hand-over over several chunks: clips per exact re-score launch (0 = per chunk)
OPT_CTC_SPLIT = 7         # wide-vocabulary CTC backward: 1 beta, rows (default); 2 beta, fill, patches; 3 beta || fill (rejected); 0 fill, recursion
NORM_STANDARD, NORM_MINMAX = 0, 1
ERR_BUSY = -6

EXPORTS = [
    "ww_version", "ww_create", "ww_destroy", "ww_last_error", "ww_load_weights", "ww_num_frames",
    "ww_mfcc_batch", "ww_cmvn", "ww_cnn_forward", "ww_quantize_weights_i8", "ww_cnn_forward_i8", "ww_score_clips", "ww_score_clips_host",
    "ww_stream_score", "ww_stream_events", "ww_session_open", "ww_session_write", "ww_session_write_tdm", "ww_session_poll",
    "ww_session_windows", "ww_session_last_logits", "ww_session_close", "ww_ctc_greedy", "ww_ctc_loss_workspace_bytes",
    "ww_ctc_loss_fwd", "ww_ctc_loss_bwd", "ww_debug_tc", "ww_debug_esp_tables", "ww_extract_mfcc", "ww_free_mfcc", "ww_analyze_mfcc_range", "ww_ring_create", "ww_ring_delete", "ww_ring_write", "ww_ring_read", "ww_ring_count",
    "ww_set_option", "ww_wav_parse", "ww_wav_load_batch", "ww_wav_write", "ww_tdm_downmix", "ww_augment_waveform",
    "ww_score_wav_files", "ww_tc_band_info", "ww_tc_rescored_total", "ww_normalize_rows", "ww_stream_score_segment", "ww_extract_mfcc_ctx",
]


class WavInfo(C.Structure):
    """ww_wav_info (include/ww_b200.h): the fields of wav::WavHeader, esp_wav.hpp:24-40."""
    _fields_ = [("riff_length", C.c_uint32), ("fmt_length", C.c_uint32), ("audio_format", C.c_uint16),
                ("num_channels", C.c_uint16), ("sample_rate", C.c_uint32), ("byte_rate", C.c_uint32),
                ("block_align", C.c_uint16), ("bits_per_sample", C.c_uint16), ("data_length", C.c_uint32),
                ("raw_data_pos", C.c_uint32), ("n_samples", C.c_uint32), ("valid", C.c_int32)]


class MfccRange(C.Structure):
    """ww_mfcc_range (include/ww_b200.h): what analyze_mfcc_range (mfcc.c:530-553) logs."""
    _fields_ = [("min_val", C.c_float), ("max_val", C.c_float), ("avg", C.c_float), ("valid", C.c_longlong),
                ("size", C.c_longlong)]


class WWError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load_library():
    """dlopen libwwb200.so and declare the prototypes.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise WWError(
                f"{LIB_PATH} is not built; run `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  ww_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float
        lib.ww_version.restype = i32
        lib.ww_create.argtypes = [C.POINTER(vp), i32]
        lib.ww_destroy.argtypes = [vp]
        lib.ww_destroy.restype = None
        lib.ww_last_error.argtypes = [vp]
        lib.ww_last_error.restype = C.c_char_p
        lib.ww_load_weights.argtypes = [vp, vp, vp, vp, vp, vp, i32]
        lib.ww_num_frames.argtypes = [i32, i32]
        lib.ww_mfcc_batch.argtypes = [vp, vp, i32, i64, i32, i64, i32, i32, vp, vp]
        lib.ww_cmvn.argtypes = [vp, vp, i64, i32, vp, vp]
        lib.ww_cnn_forward.argtypes = [vp, vp, i64, i64, i64, i64, i32, i32, f32, i32, vp, vp, vp]
        lib.ww_quantize_weights_i8.argtypes = [vp, vp]
        lib.ww_cnn_forward_i8.argtypes = [vp, vp, i64, vp, vp]
        lib.ww_score_clips.argtypes = [vp, vp, i32, i64, i32, i32, f32, i32, vp, vp, vp]
        lib.ww_score_clips_host.argtypes = [vp, vp, i32, i64, i32, i32, f32, i32, vp, vp]
        lib.ww_stream_score.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, vp]
        lib.ww_stream_events.argtypes = [vp, i64, i32, f32, i32, i32, vp, i64]
        lib.ww_stream_events.restype = i64
        lib.ww_session_open.argtypes = [vp, i32, i32, i32, i32, f32, i32, i32, C.POINTER(vp)]
        lib.ww_session_write.argtypes = [vp, vp, i32]
        lib.ww_session_write_tdm.argtypes = [vp, vp, i32]
        lib.ww_session_poll.argtypes = [vp, vp, i64]
        lib.ww_session_poll.restype = i64
        lib.ww_session_windows.argtypes = [vp]
        lib.ww_session_windows.restype = i64
        lib.ww_session_last_logits.argtypes = [vp, C.POINTER(C.POINTER(C.c_float))]
        lib.ww_session_last_logits.restype = i64
        lib.ww_session_close.argtypes = [vp]
        lib.ww_session_close.restype = None
        lib.ww_ctc_greedy.argtypes = [vp, vp, i64, i64, i32, i32, i32, vp, i32, vp, vp, vp, i32, vp, vp]
        lib.ww_ctc_loss_workspace_bytes.argtypes = [i32, i32, i32]
        lib.ww_ctc_loss_workspace_bytes.restype = C.c_size_t
        lib.ww_ctc_loss_fwd.argtypes = [vp, vp, i64, i64, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp]
        lib.ww_ctc_loss_bwd.argtypes = [vp, vp, i64, i64, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp,
                                        i64, i64, vp]
        lib.ww_debug_tc.argtypes = [vp, vp, vp]
        lib.ww_debug_esp_tables.argtypes = [vp, vp, vp, vp]
        lib.ww_debug_esp_tables.restype = C.c_uint
        lib.ww_extract_mfcc.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32]
        lib.ww_extract_mfcc.restype = C.POINTER(C.c_float)
        lib.ww_free_mfcc.argtypes = [C.POINTER(C.c_float)]
        lib.ww_free_mfcc.restype = None
        lib.ww_analyze_mfcc_range.argtypes = [vp, i64, C.c_char_p, C.POINTER(MfccRange)]
        lib.ww_analyze_mfcc_range.restype = i64
        lib.ww_ring_create.argtypes = [C.POINTER(vp), i32]
        lib.ww_ring_delete.argtypes = [vp]
        lib.ww_ring_delete.restype = None
        lib.ww_ring_write.argtypes = [vp, vp, i64]
        lib.ww_ring_read.argtypes = [vp, vp, i32]
        lib.ww_ring_count.argtypes = [vp]
        lib.ww_set_option.argtypes = [vp, i32, i32]
        lib.ww_wav_parse.argtypes = [vp, C.c_size_t, i32, C.POINTER(WavInfo)]
        lib.ww_wav_load_batch.argtypes = [C.POINTER(C.c_char_p), i32, i32, i32, vp, C.POINTER(WavInfo), C.POINTER(i32)]
        lib.ww_score_wav_files.argtypes = [vp, C.POINTER(C.c_char_p), i64, i32, i32, i32, f32, i32, vp, vp,
                                           C.POINTER(WavInfo), C.POINTER(i32), C.POINTER(C.c_double)]
        lib.ww_score_wav_files.restype = i64
        lib.ww_wav_write.argtypes = [C.c_char_p, vp, C.c_size_t, i32, i32]
        lib.ww_tdm_downmix.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp]
        lib.ww_augment_waveform.argtypes = [vp, vp, i64, i32, vp, vp]
        fp = C.POINTER(C.c_float)
        lib.ww_tc_band_info.argtypes = [vp, fp, fp, fp, fp]
        lib.ww_tc_rescored_total.argtypes = [vp, i32]
        lib.ww_tc_rescored_total.restype = i64
        lib.ww_normalize_rows.argtypes = [vp, vp, i64, i32, i64, i32, vp, vp]
        lib.ww_stream_score_segment.argtypes = [vp, vp, i32, i64, i64, i64, i64, i64, i32, i32, vp, vp, vp]
        lib.ww_extract_mfcc_ctx.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32]
        lib.ww_extract_mfcc_ctx.restype = C.POINTER(C.c_float)
        _lib = lib
        return lib


class Context:
    """One engine context per GPU (opaque ww_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.ww_create(C.byref(h), int(device))
        if rc != WW_OK or not h:
            raise WWError(f"ww_create(device={device}) failed with {rc}: no usable B200 "
                          "(ww_b200 has no CPU fallback)")
        self.h = h
        self.device = int(device)
        self.num_classes = 0

    def check(self, rc: int, what: str):
        if rc != WW_OK:
            msg = self.lib.ww_last_error(self.h)
            raise WWError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def tc_band_info(self):
        """Guard band of the tcgen05 CNN for the weights loaded last (ww_tc_band_info): |tensor logit - fp32 logit| <=
        beta * ||window||_F.  `enabled` False: cnn_impl='tensor' runs the fp32 kernel for every window."""
        v = [C.c_float() for _ in range(4)]
        rc = self.lib.ww_tc_band_info(self.h, *[C.byref(x) for x in v])
        if rc < 0:
            raise WWError(f"ww_tc_band_info failed ({rc})")
        return {"enabled": bool(rc), "beta": v[0].value, "beta_calibrated": v[1].value, "beta_rigorous": v[2].value,
                "norm_limit": v[3].value, "band_python_cmvn": v[0].value * 28.3901391,
                "band_device_cmvn": v[0].value * 42.9243521}

    def close(self):
        if getattr(self, "h", None):
            self.lib.ww_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def get_context(device=None) -> Context:
    """Process-wide context for a CUDA device index (default: torch's current device)."""
    import torch

    if device is None:
        if not torch.cuda.is_available():
            raise WWError("CUDA is not available; ww_b200 has no CPU fallback")
        device = torch.cuda.current_device()
    device = int(device)
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
