"""CPU, world_size 2, gloo: the N > 1 host logic (clip sharding + final score gather) without any GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ww_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard.shard_range(n_total, rank, world)
        # stand-in for the per-GPU scorer: a deterministic function of the global clip index
        idx = torch.arange(a, b)
        logits = (idx.float() * 0.25 - 3.0)[:, None]
        dec = (logits[:, 0] > 0).to(torch.uint8)
        g_dec = shard.gather_scores(dec, n_total)
        g_log = shard.gather_scores(logits, n_total)
        want = torch.arange(n_total).float() * 0.25 - 3.0
        ok = torch.equal(g_log[:, 0], want) and torch.equal(g_dec, (want > 0).to(torch.uint8))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11, 1])
def test_two_rank_shard_and_gather(n_total):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stream_segments_cover_windows_with_halo():
    n = 3600 * 16000
    segs = shard.stream_segments(n, 8)
    n_frames = 1 + n // 256
    assert sum(s[3] for s in segs) == n_frames - 62 == 224939
    for s0, s1, w0, nw in segs:
        # every frame of every window of the segment has its 320 taps and the pre-emphasis sample in front of them
        # (or the true stream edge, where the frontend reflects); the start keeps the TMA staging alignment
        first, last = 256 * w0 - 161, 256 * (w0 + nw - 1 + 62) + 159
        assert s0 <= max(first, 0) and s1 > min(last, n - 1) and s0 % 8 == 0
        assert s0 >= max(first, 0) - 7 and s1 <= min(last + 1, n)          # and nothing more than that
    # halo: consecutive segments share 62 frames = 61 hops + 320 taps + the pre-emphasis sample (+ < 8 of alignment)
    for a, b in zip(segs, segs[1:]):
        assert 61 * 256 + 321 <= a[1] - b[0] < 61 * 256 + 321 + 8
    # degenerate splits: more ranks than windows
    tiny = shard.stream_segments(63 * 256, 8)
    assert sum(s[3] for s in tiny) == 2 and [s[3] for s in tiny].count(0) == 6


def _ctc_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the same 7-utterance batch on every rank (seeded); each rank owns a contiguous range of utterances
        g = torch.Generator().manual_seed(5)
        T, B, C, S = 63, 7, 3, 2
        lp_all = torch.randn((T, B, C), generator=g).log_softmax(-1)
        tg_all = torch.randint(1, C, (B, S), generator=g)
        tl_all = torch.tensor([1, 2, 2, 1, 2, 1, 2])
        il_all = torch.full((B,), T)
        # single-process reference: nn.CTCLoss(reduction='mean') over the whole batch
        lp_ref = lp_all.clone().requires_grad_(True)
        want = torch.nn.functional.ctc_loss(lp_ref, tg_all, il_all, tl_all, blank=0, reduction="mean")
        want.backward()
        a, b = shard.shard_range(B, rank, world)
        lp = lp_all[:, a:b].clone().requires_grad_(True)
        nll = torch.nn.functional.ctc_loss(lp, tg_all[a:b], il_all[a:b], tl_all[a:b], blank=0, reduction="none")
        loss, mean = shard.ctc_mean_across_ranks(nll, tl_all[a:b])
        loss.backward()
        ok = torch.allclose(mean, want.detach(), rtol=1e-6, atol=1e-7)
        ok = ok and torch.allclose(lp.grad, lp_ref.grad[:, a:b], rtol=1e-5, atol=1e-7)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_ctc_mean_equals_single_process_mean():
    """8e: utterances sharded over two ranks; one 2-scalar all-reduce reproduces reduction='mean' and its gradient."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_ctc_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_ctc_mean_single_process_is_plain_mean():
    nll = torch.tensor([2.0, 3.0, 8.0], requires_grad=True)
    loss, mean = shard.ctc_mean_across_ranks(nll, [1, 0, 4])      # a zero target length is clamped like torch does
    assert torch.allclose(loss, torch.tensor((2.0 + 3.0 + 2.0) / 3)) and torch.allclose(mean, loss.detach())
