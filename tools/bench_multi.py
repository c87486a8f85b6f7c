#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] on N GPUs of one box, plus the host->device ceiling of the e2e leg.

Launch like bench.py:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                       --master-port P tools/bench_multi.py [--what streams,split,ctc,h2d] [--seconds 3600]
(N = 1 works without torchrun).  One process per GPU, NCCL only where the path has a real exchange:
  streams  N synthetic 1-hour streams, one per GPU (seed 4321 + stream id): windows/s and audio-seconds/s, no exchange
  split    ONE 1-hour stream cut into N halo segments (ww_b200.shard.stream_segments), each GPU scores its own
           (ww_stream_score_segment), logits gathered to rank 0 and compared BIT FOR BIT with the whole stream scored there
  ctc      CTC loss fwd+bwd with the CUDA kernels, utterances sharded over the ranks, reduction='mean' of the whole
           batch through shard.ctc_mean_across_ranks (one all-reduce of two scalars): loss and gradient slices equal
           to the single-process result of the same kernels and to torch.nn.functional.ctc_loss
  h2d      pinned host -> device copies, every rank alone and all ranks at once (the ceiling of bench.py's `e2e`)
Every time is CUDA-event (or wall clock for the host-timed h2d leg) max over ranks.  Prints one JSON object per line
on rank 0 (kept under profiles/).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402
from ww_b200 import shard  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))


def emit(obj):
    if rank == 0:
        print(json.dumps(obj), flush=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, dev):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, dev, reps=3, warm=2):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1) / reps * 1e-3, dev)


def synth_stream(dev, seconds, stream_id):
    """white-noise bed with a 1 s burst every 10 s, seed 4321 + stream id."""
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + stream_id)
    n = seconds * 16000
    x = torch.randn(n, generator=g, device=dev) * 0.02
    burst = torch.randn(n, generator=g, device=dev) * 0.2
    t = torch.arange(n, device=dev)
    x = x + burst * (((t // 16000) % 10) == 3)
    return torch.round(torch.clamp(x, -1, 32767 / 32768) * 32767).to(torch.int16)


def streams_leg(dev, sd, seconds):
    pcm = synth_stream(dev, seconds, rank)
    for cmvn, impl in (("device", "tensor"), ("python", "tensor"), ("device", "int8")):
        sc = ww_b200.StreamScorer(sd, device=local, cmvn=cmvn, cnn_impl=impl)
        dt = timed(lambda: sc.score(pcm), dev)
        _, logits = sc.score(pcm)
        hits = torch.tensor([len(ww_b200.events(logits))], device=dev)
        if world > 1:
            dist.all_reduce(hits)
        W = logits.shape[0]
        emit({"config": f"configs[3] streaming: {world} x {seconds} s streams, one per GPU, hop 1 frame", "n_gpus": world,
              "cmvn": cmvn, "cnn_impl": impl, "windows_per_stream": W, "seconds_per_pass_max_over_ranks": dt,
              "windows_per_s": world * W / dt, "audio_seconds_per_s": world * seconds / dt, "hits_all_streams": int(hits.item())})


def split_leg(dev, sd, seconds):
    n = seconds * 16000
    pcm = synth_stream(dev, seconds, 0)            # every rank synthesises the SAME stream (same seed) ...
    segs = shard.stream_segments(n, world)
    s0, s1, w0, nw = segs[rank]
    mine = pcm[s0:s1].clone()                      # ... and keeps only its own segment + halo
    whole_on_0 = pcm if rank == 0 else None
    del pcm
    for cmvn, impl in (("device", "tensor"), ("python", "tensor")):
        sc = ww_b200.StreamScorer(sd, device=local, cmvn=cmvn, cnn_impl=impl)
        dt = timed(lambda: sc.score_segment(mine, s0, n, w0, nw), dev)
        _, lg = sc.score_segment(mine, s0, n, w0, nw)
        # the final gather (4 B per window): pad to the widest shard
        width = max(s[3] for s in segs)
        pad = torch.zeros((width, lg.shape[1]), device=dev)
        pad[:nw] = lg
        allp = [torch.zeros_like(pad) for _ in range(world)]
        if world > 1:
            dist.all_gather(allp, pad)
        else:
            allp = [pad]
        same = None
        dt_whole = None
        if rank == 0:
            stitched = torch.cat([p[:s[3]] for p, s in zip(allp, segs)])
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            sc.score(whole_on_0)
            torch.cuda.synchronize()
            t0.record()
            _, whole = sc.score(whole_on_0)
            t1.record()
            torch.cuda.synchronize()
            dt_whole = t0.elapsed_time(t1) * 1e-3
            same = bool(torch.equal(stitched, whole))
        barrier()
        emit({"config": f"configs[3] streaming: ONE {seconds} s stream split into {world} halo segments", "n_gpus": world,
              "cmvn": cmvn, "cnn_impl": impl, "windows": sum(s[3] for s in segs), "halo_samples": 61 * 256 + 321,
              "seconds_per_pass_max_over_ranks": dt, "windows_per_s": sum(s[3] for s in segs) / dt,
              "audio_seconds_per_s": seconds / dt, "one_gpu_whole_stream_seconds": dt_whole,
              "speedup_over_one_gpu": (dt_whole / dt) if dt_whole else None,
              "stitched_logits_bit_equal_to_whole_stream": same})


def ctc_leg(dev):
    """ml_models/test.py:99-119 (T = 63, C = 3) and ml_models/ctc.py:384-407 (T = 801, wide vocabulary) training-step
    shapes, data-parallel over utterances."""
    for (T, B_per, C, S) in ((63, 1 << 18, 3, 2), (801, 256, 4096, 32)):
        B = B_per * world
        # parity on a batch every rank can hold whole: the same seeded global batch everywhere, each rank owns a range
        Bp = min(B, 64 * world if C > 64 else 4096 * world)
        g = torch.Generator(device=dev)
        g.manual_seed(777)
        lp_all = torch.log_softmax(torch.randn((T, Bp, C), generator=g, device=dev), dim=-1)
        tg_all = torch.randint(1, C, (Bp, S), generator=g, device=dev)
        tl_all = torch.randint(1, S + 1, (Bp,), generator=g, device=dev).to(torch.int32)
        il_all = torch.full((Bp,), T, dtype=torch.int32, device=dev)
        a, b = shard.shard_range(Bp, rank, world)
        lp = lp_all[:, a:b].clone().requires_grad_(True)
        nll = ww_b200.ctc_loss(lp, tg_all[a:b], il_all[a:b], tl_all[a:b], blank=0, reduction="none", zero_infinity=True)
        loss, mean = shard.ctc_mean_across_ranks(nll, tl_all[a:b])
        loss.backward()
        lp_one = lp_all.clone().requires_grad_(True)       # single-process result of the same kernels, on every rank
        one = ww_b200.ctc_loss(lp_one, tg_all, il_all, tl_all, blank=0, reduction="mean", zero_infinity=True)
        one.backward()
        lp_t = lp_all.clone().requires_grad_(True)         # and torch's own CUDA kernel
        ref = torch.nn.functional.ctc_loss(lp_t, tg_all, il_all.long(), tl_all.long(), blank=0, reduction="mean",
                                           zero_infinity=True)
        ref.backward()
        d_loss = abs(float(mean) - float(one.detach())) / max(1e-12, abs(float(one.detach())))
        d_grad = float((lp.grad - lp_one.grad[:, a:b]).abs().max())
        d_loss_t = abs(float(mean) - float(ref.detach())) / max(1e-12, abs(float(ref.detach())))
        d_grad_t = float((lp.grad - lp_t.grad[:, a:b]).abs().max())
        worst = torch.tensor([d_loss, d_grad, d_loss_t, d_grad_t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        del lp_all, lp_one, lp_t, lp
        # throughput at the full per-GPU size
        g.manual_seed(1000 + rank)
        lpb = torch.log_softmax(torch.randn((T, B_per, C), generator=g, device=dev), dim=-1)
        tgb = torch.randint(1, C, (B_per, S), generator=g, device=dev)
        ilb = torch.full((B_per,), T, dtype=torch.int32, device=dev)
        tlb = torch.full((B_per,), S, dtype=torch.int32, device=dev)

        def step():
            x = lpb.detach().requires_grad_(True)
            nl = ww_b200.ctc_loss(x, tgb, ilb, tlb, blank=0, reduction="none", zero_infinity=True)
            ls, _ = shard.ctc_mean_across_ranks(nl, tlb)
            ls.backward()

        def step_torch():
            x = lpb.detach().requires_grad_(True)
            nl = torch.nn.functional.ctc_loss(x, tgb, ilb.long(), tlb.long(), blank=0, reduction="none", zero_infinity=True)
            ls, _ = shard.ctc_mean_across_ranks(nl, tlb)
            ls.backward()

        dt = timed(step, dev)
        dt_t = timed(step_torch, dev)
        emit({"config": f"configs[4] CTC loss fwd+bwd, T={T} C={C} S={S}, {B_per} utterances per GPU, mean over the global batch",
              "n_gpus": world, "global_batch": B, "seq_per_s": B / dt, "torch_cuda_kernel_same_harness_seq_per_s": B / dt_t,
              "algorithmic_GBps_per_gpu": 2 * T * C * 4 * B_per / dt / 1e9,
              "parity_batch": Bp, "loss_rel_diff_vs_single_process": float(worst[0]),
              "grad_max_abs_diff_vs_single_process": float(worst[1]), "loss_rel_diff_vs_torch": float(worst[2]),
              "grad_max_abs_diff_vs_torch": float(worst[3]),
              "ok": bool(worst[0] < 1e-6 and worst[1] < 1e-7 and worst[2] < 1e-5 and worst[3] < 1e-5)})
        del lpb


def h2d_leg(dev, mib=2048):
    n = mib << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    d = torch.empty(n, dtype=torch.uint8, device=dev)

    def copy_rate(reps=4):
        for _ in range(2):
            d.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    alone = []
    for r in range(world):          # one rank at a time
        barrier()
        dt = copy_rate() if r == rank else 0.0
        barrier()
        alone.append(max_over_ranks(dt, dev))
    barrier()
    dt_all = max_over_ranks(copy_rate(), dev)     # all at once
    mine = torch.tensor([n / copy_rate() / 1e9], dtype=torch.float64, device=dev)
    every = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(every, mine)
    else:
        every = [mine]
    emit({"config": f"pinned host -> device copy ceiling, {mib} MiB per rank", "n_gpus": world,
          "alone_gbs_per_rank": [n / t / 1e9 for t in alone], "all_at_once_aggregate_gbs": world * n / dt_all / 1e9,
          "all_at_once_per_rank_gbs_second_pass": [float(t.item()) for t in every],
          "cpus_visible": len(os.sched_getaffinity(0)), "clips_per_s_at_that_rate": world * n / dt_all / 32000})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="h2d,streams,split,ctc")
    ap.add_argument("--seconds", type=int, default=3600)
    args = ap.parse_args()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    sd = bench.load_weights()
    for leg in args.what.split(","):
        if leg == "h2d":
            h2d_leg(dev)
        elif leg == "streams":
            streams_leg(dev, sd, args.seconds)
        elif leg == "split":
            split_leg(dev, sd, args.seconds)
        elif leg == "ctc":
            ctc_leg(dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
