#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs[3] (streaming) and configs[4] (CTC loss fwd+bwd).

Run on the GPU box; prints one JSON object per measurement (kept under profiles/).  Timing: CUDA events on
the launching stream, 3 warm-ups, inputs larger than L2 or freshly produced per iteration.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "esp32-wake-word_b200"))
import bench  # noqa: E402
import ww_b200  # noqa: E402


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def stream_bench(dev, sd, seconds=3600, impl="tensor"):
    """One 1-hour synthetic stream (seed 4321): white-noise bed + a 1 s burst every 10 s."""
    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    n = seconds * 16000
    x = torch.randn(n, generator=g, device=dev) * 0.02
    burst = torch.randn(n, generator=g, device=dev) * 0.2
    t = torch.arange(n, device=dev)
    x = x + burst * (((t // 16000) % 10) == 3)
    pcm = torch.round(torch.clamp(x, -1, 32767 / 32768) * 32767).to(torch.int16)
    del x, burst, t
    out = []
    for cmvn, impl in (("device", impl), ("python", impl), ("device", "int8")):
        sc = ww_b200.StreamScorer(sd, device=0, cmvn=cmvn, cnn_impl=impl)
        dt = timed(lambda: sc.score(pcm))
        feats, logits = sc.score(pcm)
        hits = ww_b200.events(logits)
        out.append({"config": "configs[3] streaming, 1 x 3600 s stream, hop 1 frame", "cmvn": cmvn, "cnn_impl": impl,
                    "frames": int(feats.shape[1]), "windows": int(logits.shape[0]), "seconds_per_pass": dt,
                    "audio_seconds_per_s": seconds / dt, "windows_per_s": logits.shape[0] / dt, "hits": len(hits)})
    return out


def ctc_bench(dev, T, B, C, S, seed=777, compare=True):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lp = torch.log_softmax(torch.randn((T, B, C), generator=g, device=dev), dim=-1)
    tg = torch.randint(1, C, (B, S), generator=g, device=dev)
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    crit = ww_b200.CTCLoss(blank=0, zero_infinity=True)
    ref = torch.nn.CTCLoss(blank=0, zero_infinity=True)

    def ours():
        x = lp.detach().requires_grad_(True)
        crit(x, tg, il, tl).backward()

    def torch_gpu():
        x = lp.detach().requires_grad_(True)
        ref(x, tg, il.long(), tl.long()).backward()

    d_ours = timed(ours)
    if not compare:
        return {"config": f"configs[4] CTC loss fwd+bwd T={T} B={B} C={C} S={S}", "seq_per_s": B / d_ours,
                "ms": d_ours * 1e3, "algorithmic_GBps": 2 * T * C * 4 * B / d_ours / 1e9}
    try:
        d_torch = timed(torch_gpu)
    except Exception:   # torch's CUDA kernel rejects very large batches (grid limit)
        d_torch = float("nan")
        torch.cuda.synchronize()
    nb = min(B, 4096)
    lpc, tgc = lp[:, :nb].cpu(), tg[:nb].cpu()
    t0 = time.perf_counter()
    x = lpc.detach().requires_grad_(True)
    ref(x, tgc, il[:nb].cpu().long(), tl[:nb].cpu().long()).backward()
    d_cpu = (time.perf_counter() - t0) * B / nb
    bytes_alg = 2 * T * C * 4 * B
    return {"config": f"configs[4] CTC loss fwd+bwd T={T} B={B} C={C} S={S}", "seq_per_s": B / d_ours,
            "torch_cuda_seq_per_s": B / d_torch, "torch_cpu_seq_per_s": B / d_cpu,
            "algorithmic_GBps": bytes_alg / d_ours / 1e9}


def cmvn_bench(dev, n=1 << 18):
    """a8: stand-alone CMVN (normalize_mfcc) over [n, 13, 63] windows, 6 552 B per window; and the feature-extraction
    chain mfcc_batch + normalize_mfcc that extract_features runs."""
    peak = 6545.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    feats = torch.randn((n, 13, 63), device=dev) * 10 - 20
    out = []
    for dev_style in (False, True):
        dt = timed(lambda: ww_b200.cmvn_batch(feats, device_style=dev_style))
        out.append({"config": "a8 stand-alone CMVN (ww_cmvn)", "style": "device" if dev_style else "python", "windows": n,
                    "windows_per_s": n / dt, "algorithmic_GBps": n * 6552 / dt / 1e9, "hbm_frac": n * 6552 / dt / 1e9 / peak})
    pcm = bench.synth_pcm(n, dev, 1234)
    dt = timed(lambda: ww_b200.normalize_mfcc(ww_b200.mfcc_batch(pcm), "cmvn"))
    out.append({"config": "feature extraction chain: mfcc_batch + normalize_mfcc('cmvn')", "clips": n, "clips_per_s": n / dt})
    return out


def greedy_bench(dev):
    """a11/a12: CTC best-path decode, bandwidth-bound argmax over [B, T, C] log-probs (T*C*4 bytes per utterance)."""
    peak = 6545.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = []
    for (B, T, C) in ((256, 801, 4096), (1 << 20, 63, 3), (4096, 801, 64)):
        lp = torch.randn((B, T, C), device=dev)
        dt = timed(lambda: ww_b200.greedy_batch(lp, mode="collapse"))
        t0 = time.perf_counter()
        torch.argmax(lp, dim=-1)
        torch.cuda.synchronize()
        dt_t = timed(lambda: torch.argmax(lp, dim=-1))
        gb = B * T * C * 4 / 1e9
        gbo = gb + B * (T * 4 + 4) / 1e9      # + the int32 [B, T] label matrix and the lengths it has to write
        out.append({"config": f"a12 CTC greedy decode B={B} T={T} C={C}", "seq_per_s": B / dt, "algorithmic_GBps": gb / dt,
                    "hbm_frac": gb / dt / peak, "with_outputs_GBps": gbo / dt, "hbm_frac_with_outputs": gbo / dt / peak,
                    "torch_argmax_only_GBps": gb / dt_t})
        del lp
    return out


def batch_sweep(dev, sd):
    """configs[2]: per-GPU batch-size sweep of the fused scorer (tensor CNN) and the frontend alone."""
    out = []
    big = bench.synth_pcm(1 << 20, dev, 1234)
    sc = ww_b200.WakeWordScorer(sd, device=0, cnn_impl="tensor")
    for lg in range(10, 21, 2):
        B = 1 << lg
        x = big[:B]
        dt = timed(lambda: sc.score(x), reps=5 if lg >= 16 else 20)
        df = timed(lambda: ww_b200.mfcc_batch(x) if lg <= 18 else None, reps=5 if lg >= 16 else 20) if lg <= 18 else None
        out.append({"config": "configs[2] batch sweep, fused MFCC+CMVN+CNN(tcgen05)+decision", "clips": B,
                    "ms": dt * 1e3, "clips_per_s": B / dt,
                    "frontend_alone_clips_per_s": (B / df) if df else None})
    # configs[1], fp32-input variant: 2^19 clips = 33.5 GB of float PCM
    xf = big[: 1 << 19].to(torch.float32) / 32768.0
    df = timed(lambda: ww_b200.mfcc_batch(xf), reps=3)
    out.append({"config": "configs[1] frontend alone, fp32 PCM input", "clips": 1 << 19, "clips_per_s": (1 << 19) / df,
                "algorithmic_GBps": (1 << 19) * 67276 / df / 1e9})
    del xf, big
    return out


def cpu_stages(sd):
    """configs[0]: per-stage timing of the reference's CPU path -- lives in bench.py's cpu_baseline leg (the only
    place outside tests/ that may run the oracle)."""
    return bench.cpu_baseline_stages(sd)


def session_bench(sd, n_streams=4096, chunk=320, pushes=200):
    """8f rank 2: push/poll sessions -- n_streams concurrent streams, one 20 ms chunk (320 samples) per push
    (the firmware's read_mic cadence), host PCM in, hits polled.  Wall clock (the call is synchronous)."""
    rng = np.random.default_rng(3)
    out = []
    for cmvn, impl in (("device", "tensor"), ("python", "tensor")):
        ses = ww_b200.StreamSession(sd, n_streams, max_chunk_samples=chunk, device=0, cmvn=cmvn, cnn_impl=impl)
        # pinned host memory, as a capture service would use: the H2D copy of the push is then asynchronous DMA
        data = torch.from_numpy((rng.standard_normal((n_streams, chunk)) * 600).astype(np.int16)).pin_memory()
        for _ in range(70):          # fill the 63-frame window first
            ses.write(data)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(pushes):
            ses.write(data)
            ses.poll()
        dt = (time.perf_counter() - t0) / pushes
        out.append({"config": "8f rank 2: streaming sessions, push 20 ms chunks for all streams + poll", "streams": n_streams, "host_memory": "pinned",
                    "chunk_samples": chunk, "cmvn": cmvn, "cnn_impl": impl, "ms_per_push": dt * 1e3,
                    "realtime_streams_supported": n_streams * (chunk / 16000) / dt,
                    "windows_per_s": n_streams * (chunk / 256.0) / dt})
        ses.close()
    return out


def int8_bench(dev, sd, n=1 << 18):
    """8f rank 1: int8 power-of-two twin, windows/s on the tensor cores (kind::i8) and on the CUDA cores."""
    import ctypes as C
    from ww_b200 import _lib as L
    from ww_b200.model import XIAOA_EXPONENTS, _push_weights

    x = torch.randint(-128, 128, (n, 13, 63), dtype=torch.int8, device=dev)
    ctx = L.get_context(0)
    _push_weights(ctx, sd, ("int8-bench", 0))
    exps = (C.c_int * 12)(*[int(e) for e in XIAOA_EXPONENTS])
    ctx.check(ctx.lib.ww_quantize_weights_i8(ctx.h, exps), "quantize")
    out = torch.empty((n, ctx.num_classes), dtype=torch.int8, device=dev)
    res = []
    ref = None
    for name, impl in (("tensor (tcgen05 kind::i8)", L.CNN_TENSOR), ("cuda cores", L.CNN_FP32)):
        ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_I8_IMPL, impl), "opt")
        dt = timed(lambda: ctx.check(ctx.lib.ww_cnn_forward_i8(ctx.h, L.ptr(x), n, L.ptr(out), L.cur_stream(dev)), "i8"))
        same = True if ref is None else bool(torch.equal(ref, out))
        ref = out.clone()
        res.append({"config": "8f rank 1: int8 twin forward", "impl": name, "windows": n, "seconds_per_pass": dt,
                    "windows_per_s": n / dt, "int8_TOPS_useful": n * 1291968 / dt / 1e12, "identical_to_previous_impl": same})
    ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_I8_IMPL, L.CNN_TENSOR), "opt")
    return res


def device_path_bench(dev, sd, n=1 << 18):
    """The device decision path over clips: frontend + (int8 rounding, device CMVN, int8 model, decision) in the
    kind::i8 kernel, against the float model with device CMVN."""
    pcm = bench.synth_pcm(n, dev, 1234)
    out = []
    for impl in ("int8", "tensor"):
        sc = ww_b200.WakeWordScorer(sd, device=0, cmvn="device", decision="device", cnn_impl=impl)
        dt = timed(lambda: sc.score(pcm))
        out.append({"config": "device decision path over clips (device CMVN, sigmoid*100 >= 80)", "cnn_impl": impl, "clips": n,
                    "seconds_per_pass": dt, "clips_per_s": n / dt})
    return out


def frontdsp_bench(dev):
    """SURVEY 8f rank 3/4 rows against the HBM roofline (MEASURED_PEAKS.json hbm_gbs, else 6545.3)."""
    import tempfile

    peak = 6545.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = []
    # TDM down-mix: 65 536 one-second captures = 25.2 GB in (> L2), 26 B per output sample
    B, n = 65536, 16000
    tdm = torch.randint(-32768, 32767, (B, 12 * n), dtype=torch.int16, device=dev)
    dt = timed(lambda: ww_b200.tdm_downmix(tdm))
    out.append({"config": "8f rank 4: TDM 4ch 48 kHz -> mono 16 kHz (ww_tdm_downmix)", "clips": B, "seconds_per_pass": dt,
                "clips_per_s": B / dt, "algorithmic_GBps": B * n * 26 / dt / 1e9, "hbm_frac": B * n * 26 / dt / 1e9 / peak})
    del tdm
    # augmentation: 65 536 clips, 4 B in + 20 B out per sample
    B = 65536
    x = torch.rand(B, n, device=dev) * 2 - 1
    dt = timed(lambda: ww_b200.augment_batch(x))
    out.append({"config": "8f rank 4: augment_audio_waveform x5 (ww_augment_waveform)", "clips": B, "seconds_per_pass": dt,
                "clips_per_s": B / dt, "algorithmic_GBps": B * n * 24 / dt / 1e9, "hbm_frac": B * n * 24 / dt / 1e9 / peak})
    del x
    # WAV loader (host): 4096 files of 1 s on tmpfs-or-disk, native reader threads into pinned memory
    with tempfile.TemporaryDirectory() as td:
        pcm = (np.random.default_rng(1).integers(-3000, 3000, size=16000)).astype(np.int16)
        paths = []
        for i in range(4096):
            p = os.path.join(td, f"{i}.wav")
            ww_b200.write_wav(p, pcm)
            paths.append(p)
        ww_b200.load_wav_batch(paths)
        t0 = time.perf_counter()
        ww_b200.load_wav_batch(paths)
        dt = time.perf_counter() - t0
        out.append({"config": "8f rank 3: load_wav_batch (page-cache warm, host threads)", "files": len(paths),
                    "threads": min(32, os.cpu_count() or 1), "files_per_s": len(paths) / dt, "host_GBps": len(paths) * 32044 / dt / 1e9})
    return out


def main():
    dev = torch.device("cuda", 0)
    sd = bench.load_weights()
    res = []
    if "--cpu" in sys.argv:
        for r in cpu_stages(sd):
            print(json.dumps(r), flush=True)
        return
    if "--frontdsp" in sys.argv:
        for r in cmvn_bench(dev) + greedy_bench(dev) + int8_bench(dev, sd) + device_path_bench(dev, sd) + session_bench(sd) + frontdsp_bench(dev):
            print(json.dumps(r), flush=True)
        return
    if "--greedy" in sys.argv:
        for r in greedy_bench(dev):
            print(json.dumps(r), flush=True)
        return
    if "--ctc" in sys.argv:
        for r in (ctc_bench(dev, 63, 1 << 18, 3, 2), ctc_bench(dev, 63, 1 << 18, 3, 1), ctc_bench(dev, 63, 1 << 20, 3, 2),
                  ctc_bench(dev, 801, 256, 4096, 32)):
            print(json.dumps(r), flush=True)
        return
    if "--ctc-split" in sys.argv:
        # wide-vocabulary backward, the four ways (WW_OPT_CTC_SPLIT), alternating twice
        from ww_b200 import _lib as L

        ctx = L.get_context(0)
        names = {1: "beta, rows (fill + patches)", 2: "beta, fill, patches", 3: "beta || fill, patches", 0: "fill, recursion with patches"}
        for _ in range(2):
            for mode in (1, 2, 3, 0):
                ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_CTC_SPLIT, mode), "opt")
                for shape in ((801, 256, 4096, 32), (801, 64, 4096, 32), (200, 1024, 512, 20)):
                    r = ctc_bench(dev, *shape, compare=False)
                    r["ctc_split"] = mode
                    r["how"] = names[mode]
                    print(json.dumps(r), flush=True)
        ctx.check(ctx.lib.ww_set_option(ctx.h, L.OPT_CTC_SPLIT, 1), "opt")
        # the beta recursion beside the forward pass (WW_CTC_BETA_IN_FWD, default) against the two-call form
        from ww_b200 import ctc as wctc

        for _ in range(2):
            for flag in (False, True):
                wctc.BETA_IN_FWD = flag
                for shape in ((801, 256, 4096, 32), (801, 64, 4096, 32), (200, 1024, 512, 20)):
                    r = ctc_bench(dev, *shape, compare=False)
                    r["beta_in_fwd"] = flag
                    print(json.dumps(r), flush=True)
        wctc.BETA_IN_FWD = True
        return
    res += batch_sweep(dev, sd)
    res += stream_bench(dev, sd)
    res.append(ctc_bench(dev, 63, 1 << 18, 3, 2))
    res.append(ctc_bench(dev, 63, 1 << 18, 3, 1))
    res.append(ctc_bench(dev, 801, 256, 4096, 32))
    for r in res:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
