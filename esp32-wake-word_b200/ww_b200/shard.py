"""Clip / stream sharding across the GPUs of one box.

Clips are independent (CMVN is per clip, the 161 KB of weights are replicated), so a job is split into
contiguous clip ranges, one process per GPU, with NO collective on the hot path; the only exchange is the
final gather of the per-clip scores (5 B per clip).  Streams are split into time segments that carry a halo of
62 frames (+ the centre-pad / pre-emphasis context), so segments need no exchange either.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

HOP = 256
WINDOW = 63


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [start, stop) of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scores(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather variable-length per-rank score vectors into the global order (every rank gets the result)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group) if local.is_cuda else \
        dist.all_gather(list(out.view((world, width) + tuple(local.shape[1:])).unbind(0)), pad, group=group)
    parts = [out[r * width: r * width + (b - a)] for r, (a, b) in enumerate(sizes)]
    del rank
    return torch.cat(parts, dim=0)


def stream_segments(n_samples: int, world: int):
    """Split a stream's WINDOWS into `world` contiguous ranges and give each the sample span it needs.

    Returns a list of (sample_start, sample_stop, first_window, n_windows).  Window w covers frames w..w+62 of the
    stream's frame grid; frame t is centred on sample 256 t and reads the 320 taps 256 t - 160 .. 256 t + 159 plus
    one earlier sample for the pre-emphasis (reflect padding exists only at the two true ends of the stream).  A
    segment of windows [w0, w1) therefore needs samples [256 w0 - 161, 256 (w1 + 61) + 160) clipped to the stream;
    sample_start is rounded DOWN to a multiple of 8 so that the frontend keeps its TMA staging path.  The halo
    shared by neighbours is the span of 62 frames, 61 * 256 + 321 = 15 937 samples (+ < 8 of alignment), and no
    segment ever needs data from another rank.

    A segment is consumed by `StreamScorer.score_segment(pcm[s0:s1], s0, n_samples, w0, nw)`
    (ww_stream_score_segment): it enters the frontend with the frame phase of the whole stream, so the stitched
    per-window logits are bit for bit those of `StreamScorer.score` over the whole stream.
    """
    n_frames = 1 + n_samples // HOP
    n_windows = n_frames - WINDOW + 1
    out = []
    for r in range(world):
        w0, w1 = shard_range(max(n_windows, 0), r, world)
        if w1 <= w0:
            out.append((0, 0, w0, 0))
            continue
        s0 = max(0, HOP * w0 - 161) // 8 * 8
        s1 = min(n_samples, HOP * (w1 + WINDOW - 2) + 160)
        out.append((s0, s1, w0, w1 - w0))
    return out


def ctc_mean_across_ranks(nll: torch.Tensor, target_lengths, group=None):
    """`reduction='mean'` of nn.CTCLoss (ml_models/test.py:89, ml_models/ctc.py:369: per-utterance loss / target
    length, then the batch mean) when the utterances of one batch are sharded over the ranks.

    nll: this rank's per-utterance losses [B_local] (`ctc_loss(..., reduction='none')`).  One all-reduce of two
    scalars (sum of loss/length, utterance count) is the only exchange.  Returns (loss, global_mean):
    `loss` = this rank's share, sum_local(nll_i / len_i) / N_global — differentiable, and the sum over ranks of its
    gradients (what the weight-gradient all-reduce computes) is the gradient of the global mean; `global_mean` is
    the detached value every rank reports."""
    tl = torch.as_tensor(target_lengths).to(device=nll.device, dtype=nll.dtype).clamp_min(1)
    local = (nll / tl).sum()
    stats = torch.stack([local.detach(), torch.tensor(float(nll.numel()), device=nll.device, dtype=nll.dtype)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    n_global = stats[1].clamp_min(1)
    return local / n_global, stats[0] / n_global
