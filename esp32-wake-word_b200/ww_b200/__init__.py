"""ww_b200 -- B200-native batch engine for the wake-word hot path of Socrates666/esp32-wake-word.

Mirrors the reference's `ml_models` entry points (feature, model-forward and CTC calls) on top of
libwwb200.so (hand-written sm_100a CUDA behind a C ABI, include/ww_b200.h).  No CPU fallback.
"""
from ._lib import WWError, get_context, load_library  # noqa: F401
from .features import (add_random_noise, analyze_mfcc_range, augment_audio_waveform, cmvn_batch,  # noqa: F401
                       extract_features, load_wav, mfcc_batch, normalize_mfcc, pad_audio)
from .model import LightweightKWS, WakeWordScorer, forward_int8, score_clips_int8, XIAOA_EXPONENTS  # noqa: F401
from .ctc import (CTCKeywordDetector, CTCLoss, ctc_greedy_decode, ctc_loss, decode_predictions,  # noqa: F401
                  greedy_batch)
from .stream import RingBuffer, StreamScorer, StreamSession, events, refractory_frames  # noqa: F401
from .onnx_reader import load_kws_state_dict, read_initializers  # noqa: F401
from .wav import load_wav_batch, parse_wav, read_wav, score_wav_dir, score_wav_files, write_wav  # noqa: F401
from .frontdsp import augment_batch, tdm_downmix  # noqa: F401
