"""Drop-ins for the reference's CTC calls.

  CTCLoss                      nn.CTCLoss(blank, zero_infinity) call shape of
                               ml_models/test.py:89,111-112 and ml_models/ctc.py:369,396
  ctc_greedy_decode            CTCKeywordDetector.ctc_greedy_decode, ml_models/test.py:201-217
  decode_predictions           THCHS30Trainer.decode_predictions, ml_models/ctc.py:453-471
  CTCKeywordDetector           detect / confidence logic of ml_models/test.py:158-235
                               (the GRU encoder and librosa features are out of scope; the detector here
                               consumes log-probabilities)
Batched forms return device tensors; strings are assembled on the host.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


def greedy_batch(log_probs, mode="collapse", lengths=None, keyword=None, batch_first=True):
    """Best-path labels for a batch.

    log_probs: CUDA float32 [B, T, C] (batch_first) or [T, B, C].
    mode: 'collapse' (ctc.py semantics) or 'keep_repeats' (test.py semantics: only index-0 frames dropped).
    Returns (labels int32 [B, T] zero padded, lengths int32 [B], hits uint8 [B] or None).
    """
    if not log_probs.is_cuda:
        if not torch.cuda.is_available():
            raise L.WWError("CUDA is not available; ww_b200 has no CPU fallback")
        log_probs = log_probs.cuda()
    lp = log_probs.to(torch.float32)
    if lp.stride(-1) != 1:
        lp = lp.contiguous()
    if batch_first:
        B, T, C = lp.shape
        b_stride, t_stride = lp.stride(0), lp.stride(1)
    else:
        T, B, C = lp.shape
        t_stride, b_stride = lp.stride(0), lp.stride(1)
    ctx = L.get_context(lp.device.index)
    labels = torch.empty((B, T), dtype=torch.int32, device=lp.device)
    out_len = torch.empty((B,), dtype=torch.int32, device=lp.device)
    kw = hits = None
    if keyword is not None:
        kw = torch.as_tensor(list(keyword), dtype=torch.int32, device=lp.device)
        hits = torch.empty((B,), dtype=torch.uint8, device=lp.device)
    if lengths is not None:
        lengths = lengths.to(device=lp.device, dtype=torch.int32).contiguous()
    dm = L.DECODE_COLLAPSE if mode == "collapse" else L.DECODE_KEEP_REPEATS
    ctx.check(ctx.lib.ww_ctc_greedy(ctx.h, L.ptr(lp), t_stride, b_stride, T, B, C, L.ptr(lengths), dm,
                                    L.ptr(labels), L.ptr(out_len), L.ptr(kw), 0 if kw is None else kw.numel(),
                                    L.ptr(hits), L.cur_stream(lp.device)), "ww_ctc_greedy")
    return labels, out_len, hits


def ctc_greedy_decode(log_probs, char_list):
    """ml_models/test.py:201-217: [T, C] or [1, T, C] log-probs -> str (repeats kept, index 0 dropped)."""
    lp = log_probs
    if lp.dim() == 3:
        lp = lp[0]
    labels, n, _ = greedy_batch(lp[None], mode="keep_repeats")
    ids = labels[0, : int(n[0])].tolist()
    return "".join(char_list[i] for i in ids)


def decode_predictions(log_probs, idx_to_char):
    """ml_models/ctc.py:453-471: [B, T, V] log-probs -> list[str] (textbook CTC collapse)."""
    labels, n, _ = greedy_batch(log_probs, mode="collapse")
    labels, n = labels.cpu(), n.cpu()
    return ["".join(idx_to_char.get(int(i), "<unk>") for i in labels[b, : int(n[b])]) for b in range(labels.shape[0])]


class CTCKeywordDetector:
    """Keyword scoring on log-probabilities (ml_models/test.py:158-235).

    Reference signature `CTCKeywordDetector(model, char_to_idx, keywords, threshold=0.8)`; the model-less form
    `CTCKeywordDetector(char_to_idx, keywords, threshold)` scores log-probabilities the caller already has.
    char_to_idx maps characters to class indices; a keyword hit is a substring match on the greedy decode (GPU
    kernel, repeats kept, index 0 dropped -- the reference's decoder) and its confidence is the reference's constant
    0.9, compared with `threshold` (0.8).
    """

    def __init__(self, *args, threshold=0.8, **kwargs):
        args = list(args)
        self.model = kwargs.pop("model", None)
        if args and not isinstance(args[0], dict):
            self.model = args.pop(0)                       # the reference's first positional argument
        self.char_to_idx = kwargs.pop("char_to_idx", None) if not args else args.pop(0)
        self.keywords = kwargs.pop("keywords", None) if not args else args.pop(0)
        if args:
            threshold = args.pop(0)
        if kwargs or args or self.char_to_idx is None or self.keywords is None:
            raise TypeError("CTCKeywordDetector([model,] char_to_idx, keywords, threshold=0.8)")
        self.idx_to_char = {v: k for k, v in self.char_to_idx.items()}
        self.threshold = threshold

    def ctc_greedy_decode(self, log_probs, char_list=None):
        return ctc_greedy_decode(log_probs, self.idx_to_char if char_list is None else char_list)

    @staticmethod
    def calculate_confidence(decoded_text, keyword, log_probs=None):
        return 0.9 if keyword in decoded_text else 0.0

    def extract_mfcc_features(self, waveform, n_mfcc=13):
        """test.py:218-229 takes `waveform[0]` of the chunk buffer and returns [T, 13] features.  The reference calls
        librosa there (a third feature definition, absent here and out of scope); this drop-in uses the
        engine's own frontend with the same frame geometry (n_fft 512, hop 256, 40 mels, 13 cepstra)."""
        from .features import mfcc_batch

        x = waveform[0]
        x = torch.as_tensor(x)
        if x.dim() == 2:
            x = x[0]
        return mfcc_batch(x[None])[0].transpose(0, 1).contiguous()

    def detect_keywords(self, audio_stream):
        """The reference's streaming loop (test.py:168-200): every chunk is appended to a buffer, the model scores the
        features of the buffer's OLDEST chunk (sic, :220), the greedy decode is searched for every keyword, and the
        buffer slides by five chunks.  Needs the `model` ([1, T, 13] -> log-probs [1, T, C])."""
        if self.model is None:
            raise ValueError("detect_keywords needs the model (CTCKeywordDetector(model, char_to_idx, keywords))")
        if hasattr(self.model, "eval"):
            self.model.eval()
        buffer, detected = [], []
        for chunk in audio_stream:
            buffer.append(chunk)
            if len(buffer):
                feats = self.extract_mfcc_features(buffer)
                with torch.no_grad():
                    log_probs = self.model(feats.unsqueeze(0))
                text = self.ctc_greedy_decode(log_probs.squeeze(0), self.idx_to_char)
                for kw in self.keywords:
                    if kw in text:
                        conf = self.calculate_confidence(text, kw, log_probs)
                        if conf > self.threshold:
                            detected.append((kw, conf))
                buffer = buffer[5:]
        return detected

    def detect_batch(self, log_probs):
        """log_probs [B, T, C] -> list (per utterance) of [(keyword, confidence), ...]."""
        out = [[] for _ in range(log_probs.shape[0])]
        for kw in self.keywords:
            ids = [self.char_to_idx[ch] for ch in kw]
            _, _, hits = greedy_batch(log_probs, mode="keep_repeats", keyword=ids)
            for b in torch.nonzero(hits).flatten().tolist():
                conf = 0.9
                if conf > self.threshold:
                    out[b].append((kw, conf))
        return out


# WW_CTC_BETA_IN_FWD (include/ww_b200.h): when the log-probs require a gradient the forward call also runs the beta
# recursion, beside the alpha recursion (wide vocabularies), and the backward call is the row-parallel pass alone.
# Set to False to keep the two recursions in their own calls (A/B, tests); results are identical.
BETA_IN_FWD = True
_FLAG_BETA_IN_FWD = 2


class _CTCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx_, log_probs, targets, input_lengths, target_lengths, blank, zero_infinity):
        lp = log_probs.to(torch.float32)
        if lp.stride(-1) != 1:
            lp = lp.contiguous()
        T, B, C = lp.shape
        dev = lp.device
        tg = targets.to(device=dev, dtype=torch.int32).contiguous()
        if tg.dim() != 2:
            raise ValueError("targets must be [B, S] (padded)")
        S = tg.shape[1]
        il = torch.as_tensor(input_lengths).to(device=dev, dtype=torch.int32).contiguous()
        tl = torch.as_tensor(target_lengths).to(device=dev, dtype=torch.int32).contiguous()
        eng = L.get_context(dev.index)
        ws = torch.empty(eng.lib.ww_ctc_loss_workspace_bytes(T, B, S), dtype=torch.uint8, device=dev)
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        flags = int(bool(zero_infinity))
        if BETA_IN_FWD and ctx_.needs_input_grad[0]:
            flags |= _FLAG_BETA_IN_FWD
        eng.check(eng.lib.ww_ctc_loss_fwd(eng.h, L.ptr(lp), lp.stride(0), lp.stride(1), T, B, C, L.ptr(tg), S,
                                          L.ptr(il), L.ptr(tl), int(blank), flags, L.ptr(nll),
                                          L.ptr(ws), L.cur_stream(dev)), "ww_ctc_loss_fwd")
        ctx_.save_for_backward(lp, tg, il, tl, ws)
        ctx_.blank, ctx_.zero_infinity = int(blank), flags   # the backward call gets the same flags
        return nll

    @staticmethod
    def backward(ctx_, grad_nll):
        lp, tg, il, tl, ws = ctx_.saved_tensors
        T, B, C = lp.shape
        dev = lp.device
        eng = L.get_context(dev.index)
        go = grad_nll.to(torch.float32).contiguous()
        grad = torch.empty((T, B, C), dtype=torch.float32, device=dev)
        eng.check(eng.lib.ww_ctc_loss_bwd(eng.h, L.ptr(lp), lp.stride(0), lp.stride(1), T, B, C, L.ptr(tg),
                                          tg.shape[1], L.ptr(il), L.ptr(tl), ctx_.blank, ctx_.zero_infinity,
                                          L.ptr(go), L.ptr(ws), L.ptr(grad), grad.stride(0), grad.stride(1),
                                          L.cur_stream(dev)), "ww_ctc_loss_bwd")
        return grad, None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False):
    """torch.nn.functional.ctc_loss semantics on CUDA log-probs [T, B, C] and padded targets [B, S]."""
    if not log_probs.is_cuda:
        raise L.WWError("ctc_loss expects CUDA log-probs; ww_b200 has no CPU fallback")
    nll = _CTCLossFn.apply(log_probs, targets, input_lengths, target_lengths, blank, zero_infinity)
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    tl = torch.as_tensor(target_lengths).to(device=nll.device, dtype=nll.dtype).clamp_min(1)
    return (nll / tl).mean()


class CTCLoss(nn.Module):
    """nn.CTCLoss(blank=0, reduction='mean', zero_infinity=False) call shape."""

    def __init__(self, blank=0, reduction="mean", zero_infinity=False):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                        self.zero_infinity)
