#!/usr/bin/env python
"""bench.py -- clips/s (1 s, 16 kHz) through MFCC + CMVN + CNN + decision + CTC best path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2] at its single-GPU size; configs[1] is the `roofline` kernel):
  * synthetic int16 PCM, `--clips` 1 s clips per GPU resident in HBM (four value
    distributions by clip index mod 4, seed 1234 + rank), larger than L2 so no flush is needed
  * step = ww_score_clips over all clips (frontend + CMVN + xiaoa CNN + sigmoid>0.5 decision) followed by
    the CTC best-path / keyword kernel over the binary posteriors grouped 63 windows per utterance,
    then (N > 1) a gather of the decisions to rank 0 -- the only collective
  * `value` = clips of all ranks / max-over-ranks device time (CUDA events on the launching stream)
  * `e2e`   = the same call with HOST buffers (pinned), H2D and D2H inside the timed region
  * `roofline` = the frontend kernel alone (ww_mfcc_batch over the same clips): algorithmic bytes
    35 276 B/clip over its CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth
  * `cpu_baseline` / `--impl reference`: the oracle port of the reference's CPU path (torchaudio T.MFCC +
    normalize_mfcc + LightweightKWS + sigmoid>0.5 + greedy decode) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "esp32-wake-word_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "clips/sec (1 s 16 kHz) MFCC+CNN+CTC"
UNIT = "clips/s"
FRONTEND_BYTES_PER_CLIP = 16000 * 2 + 13 * 63 * 4  # 35 276
FUSED_BYTES_PER_CLIP = 16000 * 2 + 5
# dram__bytes_read.sum + dram__bytes_write.sum of mfcc_kernel<int16, PY> per clip, from the `ncu --set full` capture
# profiles/r2h_ncu_mfcc_kernel.txt (2.1033 GB + 159.5 MB over a 65 536-clip launch; the last output lines are still in
# L2 when the launch ends -- a capture of the round's first half read 2.0974 GB + 209.5 MB): traffic == algorithmic bytes
FRONTEND_DRAM_BYTES_PER_CLIP_NCU = (2.103297e9 + 159.499264e6) / 65536
# executed warp instructions and shared-memory wavefronts per clip of the same kernel (ncu, same capture family;
# profiles/r2h_ncu_mfcc_kernel.txt, the final kernel of round 2): the issue ports (4 warp-instructions/clk/SM) and the
# shared-memory pipe (1 wavefront/clk/SM) are what the kernel is actually limited by, so their fractions are reported
# beside the HBM one
FRONTEND_WARP_INSTR_PER_CLIP = 1380660280 / 65536
FRONTEND_SMEM_WAVEFRONTS_PER_CLIP = 392761723 / 65536
# FP32-pipe cycles per clip of the same capture (profiles/r2h_mfcc_sass_summary.txt): a packed f32x2 instruction
# (FFMA2 224.4 M + FADD2 184.5 M + FMUL2 120.9 M warp instructions) holds a scheduler's FMA pipe for two cycles, a scalar
# FFMA / FMUL / FADD (106.0 + 17.3 + 17.0 M) for one; an SM has four such pipes
FRONTEND_FMA_PIPE_CYCLES_PER_CLIP = (2 * (224395264 + 184549376 + 120872064) + 106037248 + 17301504 + 17039360) / 65536
UTT = 63  # windows per CTC utterance
CPU_BATCH = 200
PARITY_DISTINCT = 65536   # exact-match check on >= 65 536 distinct clips per GPU
PARITY_STRIDED = 4096     # strided subset of the timed clips
PARITY_MARGIN = 2e-3      # |oracle logit - threshold| below which a float pipeline cannot promise the oracle's decision


def load_weights():
    d = np.load(os.path.join(ROOT, "tests", "golden", "xiaoa_weights.npz"))
    return {k: d[k] for k in d.files}


_SM_MHZ = [None]


def measured_sm_mhz():
    """SM clock sampled under load during the timed region (set by run_ours before the roofline leg)."""
    return _SM_MHZ[0]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
# synthetic PCM (same four distributions on either device; RNG streams differ between CPU and CUDA)
# ------------------------------------------------------------------------------------------------
def synth_pcm(n_clips, device, seed, chunk=32768, pin=False):
    import torch

    out = torch.empty((n_clips, 16000), dtype=torch.int16, device=device,
                      pin_memory=(pin and device == "cpu"))
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(16000, device=device, dtype=torch.float32) / 16000.0
    for c0 in range(0, n_clips, chunk):
        n = min(chunk, n_clips - c0)
        x = torch.randn((n, 16000), generator=g, device=device) * 0.1               # (0) white
        idx = torch.arange(c0, c0 + n, device=device) % 4
        u = torch.rand((n, 16000), generator=g, device=device) - 0.5                   # (1) uniform
        ph = torch.rand((n, 2), generator=g, device=device) * (2 * np.pi)
        tone = 0.3 * torch.sin(2 * np.pi * 440.0 * t[None] + ph[:, :1]) + \
            0.3 * torch.sin(2 * np.pi * 3000.0 * t[None] + ph[:, 1:]) + x * 0.1       # (2) tones + N(0,.01^2)
        sp = x.clone()
        sp[:, 9000:] = 0.0                                                             # (3) burst then silence
        y = torch.where((idx == 0)[:, None], x, torch.where((idx == 1)[:, None], u,
                        torch.where((idx == 2)[:, None], tone, sp)))
        y = torch.clamp(y, -1.0, 32767.0 / 32768.0)
        out[c0:c0 + n] = torch.round(y * 32767.0).to(torch.int16)
        del x, u, tone, sp, y
    return out


# ------------------------------------------------------------------------------------------------
# host placement for the e2e leg at N > 1: each rank's pinned buffers on the NUMA node its GPU hangs off
# ------------------------------------------------------------------------------------------------
def bind_host_to_gpu_node(local):
    """Prefer host memory (and CPUs, when the process is allowed any there) of the GPU's own NUMA node, so that the
    eight H2D streams of an eight-rank run do not all read one socket's DRAM across the inter-socket link.
    Everything here is best effort: a container without the topology files, or without permission for
    set_mempolicy, leaves the process as it was.  Returns what was done, for the JSON line."""
    import ctypes
    import torch

    info = {"node": None, "mempolicy": "unchanged", "cpus": len(os.sched_getaffinity(0))}
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        info["node"] = node
        if node < 0:
            return info
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238            # x86_64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
            info["mempolicy"] = "preferred" if rc == 0 else "errno %d" % ctypes.get_errno()
        except Exception as e:  # noqa: BLE001
            info["mempolicy"] = "failed: %s" % type(e).__name__
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        mine = cpus & os.sched_getaffinity(0)
        if mine:
            os.sched_setaffinity(0, mine)
            info["cpus"] = len(mine)
    except Exception as e:  # noqa: BLE001
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nv_rows, self._halt = [], threading.Event()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        # NVML directly as well (30 ms period; 10 ms costs the timed region about 1 %): nvidia-smi's loop gives only a couple of rows inside a 0.2 s region
        self.nv_rows, self._halt = [], threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._halt.is_set():
                    try:
                        self.nv_rows.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx,
                                             int(get_reasons(h))))
                    except Exception:
                        return
                    time.sleep(0.03)

            threading.Thread(target=poll, daemon=True).start()
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.15:
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        self._halt.set()
        # NVML reason bits (nvml.h): sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        for ts, c, m, mask in list(self.nv_rows):
            if ts < t0 or ts > t1:
                continue
            sm.append(float(c))
            mx.append(float(m))
            reasons.update(nm for nm, b in bits.items() if mask & b)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference path (oracle port) -- the checker used as the timed baseline, never as the product
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(pcm_i16, sd):
    """One pass of the reference's CPU path over a batch of clips; returns (#positives, #keyword hits)."""
    import torch

    from oracle import cnn as ocnn
    from oracle import ctc as octc
    from oracle import mfcc as omfcc

    # batches of 200 clips: the reference script's own batch size (ml_models/main.py:141) and the CPU's sweet spot
    # (profiles/r1_cpu_baseline_stages.jsonl: 22 k clips/s at B=200 vs 17 k at B=16384 on 16 threads)
    parts = []
    for c0 in range(0, pcm_i16.shape[0], CPU_BATCH):
        x = pcm_i16[c0:c0 + CPU_BATCH].to(torch.float32) / 32768.0  # torchaudio.load normalisation
        feats = omfcc.mfcc_torchaudio(x)                            # preemphasis + T.MFCC
        z = omfcc.normalize_mfcc(feats, "cmvn")
        parts.append(ocnn.forward_torch(z.numpy(), sd)[:, 0])
    logits = np.concatenate(parts)
    dec = ocnn.decide_python(logits)
    n_utt = len(logits) // UTT
    hits = 0
    if n_utt:
        lg = torch.from_numpy(logits[: n_utt * UTT].reshape(n_utt, UTT))
        lp = torch.stack([torch.nn.functional.logsigmoid(-lg), torch.nn.functional.logsigmoid(lg)], dim=-1).numpy()
        for b in range(n_utt):
            hits += octc.keyword_hit(octc.greedy_labels(lp[b], octc.MODE_COLLAPSE), [1])
    return int(dec.sum()), int(hits)


def cpu_baseline_stages(sd):
    """configs[0]: the reference's CPU path per stage (oracle port), B in {1, 200, 4096, 16384}, all threads and 1."""
    import torch

    from oracle import cnn as ocnn
    from oracle import mfcc as omfcc

    out = []
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count()
    for threads in (ncpu, 1):
        torch.set_num_threads(threads)
        for B in (1, 200, 4096, 16384, 65536):
            if threads == 1 and B > 4096:
                continue
            pcm = synth_pcm(B, "cpu", 1234, chunk=4096)
            x = pcm.to(torch.float32) / 32768.0

            def best(fn, n=5):
                fn(); fn()
                ts = []
                for _ in range(n):
                    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
                return min(ts), float(np.median(ts))

            feats = omfcc.mfcc_torchaudio(x)
            z = omfcc.normalize_mfcc(feats, "cmvn")
            zn = z.numpy()
            t_m = best(lambda: omfcc.mfcc_torchaudio(x))
            t_c = best(lambda: ocnn.forward_torch(omfcc.normalize_mfcc(feats, "cmvn").numpy(), sd))
            t_e = best(lambda: cpu_reference_step(pcm, sd))
            out.append({"config": "configs[0] reference CPU path (oracle port), per stage", "threads": threads,
                        "cpus": ncpu, "clips": B, "mfcc_clips_per_s_best": B / t_m[0], "mfcc_clips_per_s_median": B / t_m[1],
                        "cmvn_cnn_clips_per_s_best": B / t_c[0], "end_to_end_clips_per_s_best": B / t_e[0],
                        "end_to_end_clips_per_s_median": B / t_e[1]})
            del zn
    return out


_CPU_SPREAD = []   # clips/s of every timed pass of the last time_cpu call (the value reported is clips / mean pass time)


def time_cpu(clips, steps, warmup, seed=1234):
    import torch

    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1, which would cripple the baseline)
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    sd = load_weights()
    pcm = synth_pcm(clips, "cpu", seed, chunk=4096)
    for _ in range(warmup):
        cpu_reference_step(pcm, sd)
    per_pass = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        cpu_reference_step(pcm, sd)
        per_pass.append(time.perf_counter() - t0)
    dt = sum(per_pass) / len(per_pass)
    _CPU_SPREAD[:] = [clips / t for t in per_pass]
    return clips / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = args.cpu_clips
    v, dt, cores = time_cpu(clips, args.steps, args.warmup)
    sample = f"{clips} synthetic clips per step (bounded sample of the {args.clips}-clip/GPU workload), " \
             f"torch {cores} threads of {os.cpu_count()} cpus"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "per_pass": list(_CPU_SPREAD)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def parity_check(ctx, L, pcm_dev, logits_dev, dec_dev, idx, sd, sp, threads):
    """The oracle (torchaudio MFCC + normalize_mfcc + LightweightKWS, the reference's own calls) over clips `idx` of the
    TIMED device PCM, against what the timed step left in `logits_dev` / `dec_dev` for those clips, and against the
    frontend's features for them.  Checker only: nothing here is timed or shipped."""
    import torch

    from oracle import cnn as ocnn
    from oracle import mfcc as omfcc

    torch.set_num_threads(max(1, threads))
    sub = pcm_dev[idx].contiguous()
    n = sub.shape[0]
    feats = torch.empty((n, 13, 63), dtype=torch.float32, device=pcm_dev.device)
    ctx.check(ctx.lib.ww_mfcc_batch(ctx.h, L.ptr(sub), L.PCM_S16, n, 16000, 16000, L.FEAT_PY, L.LAYOUT_COEF_MAJOR,
                                    L.ptr(feats), sp), "ww_mfcc_batch(parity)")
    got_f = feats.cpu().numpy()
    got_l = logits_dev[idx, 0].cpu().numpy()
    got_d = dec_dev[idx].cpu().numpy().astype(bool)
    host = sub.cpu()
    f_err, l_err, mism, mism_out, near = 0.0, 0.0, 0, 0, 0
    for c0 in range(0, n, 4096):
        x = host[c0:c0 + 4096].to(torch.float32) / 32768.0
        f = omfcc.mfcc_torchaudio(x)
        z = omfcc.normalize_mfcc(f, "cmvn")
        want = ocnn.forward_torch(z.numpy(), sd)[:, 0]
        wd = ocnn.decide_python(want)
        f_err = max(f_err, float(np.abs(f.numpy() - got_f[c0:c0 + 4096]).max()))
        l_err = max(l_err, float(np.abs(want - got_l[c0:c0 + 4096]).max()))
        bad = wd != got_d[c0:c0 + 4096]
        clear = np.abs(want) > PARITY_MARGIN
        mism += int(bad.sum())
        mism_out += int((bad & clear).sum())
        near += int((~clear).sum())
    return {"n": n, "feature_max_abs": f_err, "logit_max_abs": l_err, "decision_mismatches": mism,
            "decision_mismatches_outside_margin": mism_out, "near_threshold": near}


def workload_config(args, n):
    return {"workload": "configs[2]: fused MFCC(esp_mfcc/ml_models params: 320/256/512, 40 mel, 13 cep) + CMVN + "
                        "xiaoa.onnx LightweightKWS CNN + sigmoid>0.5 decision + CTC best-path/keyword over 63-window "
                        "utterances, clip-sharded",
            "clips_per_gpu": args.clips, "global_clips": args.clips * n, "clip_samples": 16000, "pcm": "int16",
            "cnn_impl": args.cnn + (" (tcgen05 kind::f16, fp32 accumulate in TMEM; clips inside the guard band calibrated "
                                    "for the loaded weights are re-scored by the fp32 kernel)" if args.cnn == "tensor" else ""),
            "parallelism": f"dp{n} (clip shards, no hot-path collective)",
            "rank_to_gpu": "rank i -> GPU i * (visible GPUs // ranks): ranks spread over the box's host bridges (e2e leg); "
                           "identity when ranks == visible GPUs",
            "l2_policy": "inputs (32 KB/clip x clips) exceed L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import ww_b200
    from ww_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (ww_b200 has no CPU fallback)")
    # Which GPU a rank takes when the box shows more GPUs than ranks: spread the ranks over the box (stride
    # n_visible / world) instead of packing them onto GPUs 0..N-1.  Measured on this pool (profiles/r2_h2d_ceiling.md):
    # the eight GPUs hang off TWO host bridges whose aggregate host->device rates saturate at ~115 and ~140 GB/s, so
    # four ranks packed onto GPUs 0-3 share one of them (28.8 GB/s each) while GPUs 0,2,4,6 get a full link each.
    # Device-side numbers do not depend on the choice; WW_NO_SPREAD=1 restores rank i -> GPU i.
    n_vis = torch.cuda.device_count()
    gpu = local
    if world > 1 and n_vis > world and n_vis % world == 0 and not os.environ.get("WW_NO_SPREAD") \
            and int(os.environ.get("LOCAL_WORLD_SIZE", world)) == world:
        gpu = local * (n_vis // world)
    local_rank, local = local, gpu
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_numa = bind_host_to_gpu_node(local) if world > 1 and not os.environ.get("WW_NO_NUMA_BIND") else None
    # stdout carries the ONE JSON line and nothing else: for the rest of the run fd 1 points at stderr (NCCL's INFO
    # lines, library banners), and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # the communicator's own account of itself (ranks, transport, NVLS) goes to stderr so that it can be checked
        os.environ["NCCL_DEBUG"] = os.environ.get("WW_NCCL_DEBUG", "INFO")
        os.environ["NCCL_DEBUG_SUBSYS"] = os.environ.get("WW_NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)  # creates the communicator
        torch.cuda.synchronize()
    sd = load_weights()
    cnn_impl = args.cnn
    scorer = ww_b200.WakeWordScorer(sd, device=local, cmvn="python", decision="python", cnn_impl=cnn_impl)
    ctx = scorer.ctx
    B = args.clips
    pcm = synth_pcm(B, dev, 1234 + rank)
    logits = torch.empty((B, 1), dtype=torch.float32, device=dev)
    dec = torch.empty((B,), dtype=torch.uint8, device=dev)
    n_utt = B // UTT
    labels = torch.empty((max(n_utt, 1), UTT), dtype=torch.int32, device=dev)
    lab_len = torch.empty((max(n_utt, 1),), dtype=torch.int32, device=dev)
    hits = torch.empty((max(n_utt, 1),), dtype=torch.uint8, device=dev)
    kw = torch.tensor([1], dtype=torch.int32, device=dev)
    gathered = torch.empty((world * B,), dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev)
    sp = L.cur_stream(dev)
    impl_id = L.CNN_TENSOR if cnn_impl == "tensor" else L.CNN_FP32

    def step():
        ctx.check(ctx.lib.ww_score_clips(ctx.h, L.ptr(pcm), L.PCM_S16, B, L.CMVN_PY, L.DECIDE_LOGIT, 0.0, impl_id,
                                         L.ptr(logits), L.ptr(dec), sp), "ww_score_clips")
        if n_utt:
            ctx.check(ctx.lib.ww_ctc_greedy(ctx.h, L.ptr(logits), 1, UTT, UTT, n_utt, 1, None, L.DECODE_COLLAPSE,
                                            L.ptr(labels), L.ptr(lab_len), L.ptr(kw), 1, L.ptr(hits), sp),
                      "ww_ctc_greedy")
        if world > 1:
            dist.all_gather_into_tensor(gathered, dec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_max = float(tt.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if clocks and clocks.get("sm_mhz"):
        _SM_MHZ[0] = clocks["sm_mhz"]
    ms_per_step = ms_max / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- roofline: the frontend kernel alone over the same clips (configs[1]) ----------------------
    roof = None
    if rank == 0:
        rb = min(B, args.roofline_clips)
        feats = torch.empty((rb, 13, 63), dtype=torch.float32, device=dev)

        def front():
            ctx.check(ctx.lib.ww_mfcc_batch(ctx.h, L.ptr(pcm), L.PCM_S16, rb, 16000, 16000, L.FEAT_PY,
                                            L.LAYOUT_COEF_MAJOR, L.ptr(feats), sp), "ww_mfcc_batch")

        for _ in range(3):
            front()
        torch.cuda.synchronize()
        reps = max(3, args.steps)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(reps):
            front()
        f1.record(stream)
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / reps
        peak, how = measured_peaks()
        ach = rb * FRONTEND_BYTES_PER_CLIP / (fms * 1e-3) / 1e9
        sm_clk = 1e6 * (measured_sm_mhz() or 1965.0)
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        roof = {"bound": "hbm", "kernel": "mfcc_kernel<int16, PY> (frontend alone, 1 persistent launch over %d clips)" % rb,
                "achieved": ach, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": ach / peak,
                "traffic": FRONTEND_DRAM_BYTES_PER_CLIP_NCU * rb, "traffic_source": "profiles/r2h_ncu_mfcc_kernel.txt "
                "(ncu --set full, per-clip DRAM bytes x clips of this launch)", "algorithmic_bytes": rb * FRONTEND_BYTES_PER_CLIP,
                "ms_per_launch": fms, "clips_per_s": rb / (fms * 1e-3),
                "algorithmic_bytes_per_clip": FRONTEND_BYTES_PER_CLIP,
                # the kernel's real limiters, same launch: per-clip counts from the ncu capture x live clips/s over the
                # per-SM peak at the SM clock sampled under load
                "other_bounds": [
                    {"bound": "issue", "achieved": FRONTEND_WARP_INSTR_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk),
                     "peak": 4.0, "unit": "warp-instr/clk/SM",
                     "frac": FRONTEND_WARP_INSTR_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk) / 4.0,
                     "per_clip": FRONTEND_WARP_INSTR_PER_CLIP, "source": "smsp__inst_executed.sum / clips (ncu)"},
                    {"bound": "shared-memory pipe", "achieved": FRONTEND_SMEM_WAVEFRONTS_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk),
                     "peak": 1.0, "unit": "wavefronts/clk/SM",
                     "frac": FRONTEND_SMEM_WAVEFRONTS_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk),
                     "per_clip": FRONTEND_SMEM_WAVEFRONTS_PER_CLIP,
                     "source": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum / clips (ncu)"},
                    {"bound": "fp32 pipes", "achieved": FRONTEND_FMA_PIPE_CYCLES_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk),
                     "peak": 4.0, "unit": "pipe-cycles/clk/SM",
                     "frac": FRONTEND_FMA_PIPE_CYCLES_PER_CLIP * rb / (fms * 1e-3) / (n_sm * sm_clk) / 4.0,
                     "per_clip": FRONTEND_FMA_PIPE_CYCLES_PER_CLIP,
                     "source": "2 x packed f32x2 + scalar FP32 warp instructions / clips (ncu source view)"}],
                "sm_clock_hz_used": sm_clk}
        del feats

    # ---- the three feature hand-overs of ww_score_clips, same clips, same bits (rank 0 of a 1-GPU run) --------------
    # default: 131 072-clip chunks through a scratch in HBM; l2_chunks: 14 208-clip frontend + CNN launch pairs (a whole
    # number of waves for both kernels) whose features stay in L2 (programmatic dependent launch, one exact re-score launch per 131 072 clips); one_kernel: a single
    # persistent launch, frontend pipelines and tcgen05 CNN groups on disjoint SMs, L2-resident ring (csrc/ww_fused.cuh)
    handoff = None
    if rank == 0 and world == 1 and cnn_impl == "tensor" and not args.no_handoff:
        hb = min(B, 1 << 18)
        lg2 = torch.empty((hb, 1), dtype=torch.float32, device=dev)
        dc2 = torch.empty((hb,), dtype=torch.uint8, device=dev)

        def score(lg, dc):
            ctx.check(ctx.lib.ww_score_clips(ctx.h, L.ptr(pcm), L.PCM_S16, hb, L.CMVN_PY, L.DECIDE_LOGIT, 0.0, impl_id,
                                             L.ptr(lg), L.ptr(dc), sp), "ww_score_clips")

        modes = [("default", [(L.OPT_FUSED, 0), (L.OPT_L2_CHUNK_CLIPS, 0)], 38758.6, "profiles/r2_fused_dram_by_chunk.txt", 3 * 131072),
                 ("l2_chunks", [(L.OPT_FUSED, 0), (L.OPT_L2_CHUNK_CLIPS, 14208)], 32123.8, "profiles/r2_handoff_dram_modes.txt", None),
                 ("one_kernel", [(L.OPT_FUSED, 2), (L.OPT_L2_CHUNK_CLIPS, 0)], 31951.2, "profiles/r2_handoff_dram_modes.txt", None)]
        handoff = {"clips": hb, "algorithmic_bytes_per_clip": FUSED_BYTES_PER_CLIP, "modes": {}}
        ref_out = None
        for name, opts, dram, src, _ in modes:
            for o, v in opts:
                ctx.check(ctx.lib.ww_set_option(ctx.h, o, v), "ww_set_option")
            for _ in range(2):
                score(lg2, dc2)
            torch.cuda.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record(stream)
            for _ in range(3):
                score(lg2, dc2)
            h1.record(stream)
            torch.cuda.synchronize()
            hms = h0.elapsed_time(h1) / 3
            if ref_out is None:
                ref_out = (lg2.clone(), dc2.clone())
            launches_1m = {"default": 8 * 2 + 1, "l2_chunks": 74 * 2 + 8, "one_kernel": 8 * 2}[name]
            handoff["modes"][name] = {"clips_per_s": hb / (hms * 1e-3), "dram_bytes_per_clip_ncu": dram,
                                      "dram_over_algorithmic": dram / FUSED_BYTES_PER_CLIP, "dram_source": src,
                                      "kernel_launches_per_2^20_clips": launches_1m,
                                      "identical_to_default": bool(torch.equal(lg2, ref_out[0]) and torch.equal(dc2, ref_out[1]))}
        for o in (L.OPT_FUSED, L.OPT_L2_CHUNK_CLIPS):
            ctx.check(ctx.lib.ww_set_option(ctx.h, o, 0), "ww_set_option")
        del lg2, dc2, ref_out

    # ---- e2e: host buffers through the public call ---------------------------------------------------
    eb = min(args.e2e_clips, B)
    host = synth_pcm(eb, "cpu", 99 + rank, chunk=8192, pin=True)
    lh = np.empty((eb, 1), np.float32)
    dh = np.empty((eb,), np.uint8)
    scorer.score_host(host, lh, dh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scorer.score_host(host, lh, dh)
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    per_rank_e2e = [te.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_e2e, te)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    per_rank_e2e = [float(t.item()) for t in per_rank_e2e]
    e2e_val = world * eb / float(te.item())
    # the ceiling of that leg: the same pinned buffer copied host -> device by every rank at once, no kernels
    d_sink = torch.empty_like(host, device=dev)
    for _ in range(2):
        d_sink.copy_(host, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(3, args.steps)):
        d_sink.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    h2d_dt = (time.perf_counter() - t0) / max(3, args.steps)
    th = torch.tensor([h2d_dt], dtype=torch.float64, device=dev)
    per_rank_h2d = [th.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_h2d, th)
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    per_rank_h2d = [eb * 32000 / float(t.item()) / 1e9 for t in per_rank_h2d]
    h2d_ceiling_gbs = world * eb * 32000 / float(th.item()) / 1e9
    del d_sink
    if host_numa is not None:
        # every rank's placement: [NUMA node of its GPU (-1 unknown), memory policy set, CPUs it may run on]
        mine = torch.tensor([-1 if host_numa["node"] is None else host_numa["node"],
                             1 if host_numa["mempolicy"] == "preferred" else 0, host_numa["cpus"]], device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        host_numa = dict(host_numa, per_rank=[[int(v) for v in t.tolist()] for t in every])

    # ---- parity of the measured run (outside every timed region) --------------------------------------
    # `logits` / `dec` still hold what the LAST timed step wrote for this rank's B clips.  Checked against the oracle:
    # a strided subset over all of them and the first PARITY_DISTINCT clips (every synthetic clip is distinct)
    parity = None
    if not args.no_parity:
        try:
            ncpu = len(os.sched_getaffinity(0))
        except Exception:
            ncpu = os.cpu_count() or 1
        rescored = int(ctx.lib.ww_tc_rescored_total(ctx.h, 0)) if cnn_impl == "tensor" else 0
        stride = max(1, B // PARITY_STRIDED)
        idx_s = torch.arange(0, B, stride, device=dev)[:PARITY_STRIDED]
        idx_d = torch.arange(0, min(B, args.parity_clips), device=dev)
        ps = parity_check(ctx, L, pcm, logits, dec, idx_s, sd, sp, ncpu // world)
        pd = parity_check(ctx, L, pcm, logits, dec, idx_d, sd, sp, ncpu // world)
        # the e2e leg's own outputs, on the host clips it scored
        ne = min(eb, 2048)
        pe_l = torch.from_numpy(lh[:ne].copy()).to(dev)
        pe_d = torch.from_numpy(dh[:ne].copy()).to(dev)
        pe = parity_check(ctx, L, host[:ne].to(dev), pe_l, pe_d, torch.arange(ne, device=dev), sd, sp, ncpu // world)
        mine = torch.tensor([ps["n"] + pd["n"] + pe["n"],
                             ps["decision_mismatches"] + pd["decision_mismatches"] + pe["decision_mismatches"],
                             ps["decision_mismatches_outside_margin"] + pd["decision_mismatches_outside_margin"] +
                             pe["decision_mismatches_outside_margin"],
                             ps["near_threshold"] + pd["near_threshold"] + pe["near_threshold"], rescored],
                            dtype=torch.float64, device=dev)
        mx = torch.tensor([max(ps["feature_max_abs"], pd["feature_max_abs"], pe["feature_max_abs"]),
                           max(ps["logit_max_abs"], pd["logit_max_abs"], pe["logit_max_abs"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        band = ctx.tc_band_info() if cnn_impl == "tensor" else None
        parity = {"oracle": "oracle/mfcc.py mfcc_torchaudio + normalize_mfcc('cmvn') + oracle/cnn.py forward_torch "
                            "(the reference's torchaudio / torch calls), fp32 on the host",
                  "what": "clips of the TIMED device PCM against the logits / decisions the last timed step wrote for them "
                          f"(per rank: {ps['n']} strided over all clips + the first {pd['n']} clips, all distinct) and "
                          f"{pe['n']} clips of the e2e leg against its host outputs",
                  "n_checked": int(mine[0].item()), "feature_max_abs": float(mx[0].item()), "feature_tolerance": 1e-3,
                  "logit_max_abs": float(mx[1].item()),
                  "decision_mismatches": int(mine[1].item()),
                  "decision_mismatches_outside_margin": int(mine[2].item()),
                  "near_threshold_excluded": int(mine[3].item()), "margin": PARITY_MARGIN,
                  "rescored": int(mine[4].item()),
                  "rescored_note": "windows the tcgen05 path handed to the exact fp32 kernel over ALL steps of this run "
                                   "(warm-up, timed, e2e), summed over ranks",
                  "tc_band": band,
                  "ok": bool(mine[2].item() == 0 and mx[0].item() < 1e-3)}

    if rank == 0:
        chunk = int(os.environ.get("WW_CHUNK_CLIPS", "131072"))  # ww_api.cu kScratchClips: clips per fused frontend + CNN pair
        chunks = (B + chunk - 1) // chunk
        # frontend + CNN per chunk; tensor path: + the fp32 re-score of the borderline clips, ONE launch per 2^20 clips when the
        # call spans several chunks (ww_api.cu kRescoreWindowClips, WW_RESCORE_WINDOW_CLIPS=0: one per chunk)
        window = int(os.environ.get("WW_RESCORE_WINDOW_CLIPS", str(1 << 20)))
        rescores = 0
        if cnn_impl == "tensor":
            rescores = chunks if (chunks == 1 or window == 0) else (B + max(window, chunk) - 1) // max(window, chunk)
        launches = args.steps * (chunks * 2 + rescores + (1 if n_utt else 0))
        cpu = None
        if world == 1 and not args.no_cpu:
            v, dt, cores = time_cpu(args.cpu_clips, 5, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "per_pass": list(_CPU_SPREAD),
                   "spread": [min(_CPU_SPREAD), max(_CPU_SPREAD)],
                   "sample": f"{args.cpu_clips} synthetic clips x 5 passes (1 warm-up) through the oracle port of the reference CPU "
                             f"path (torchaudio MFCC + CMVN + LightweightKWS + decision + greedy), {dt:.2f} s/pass"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": eb * 32000, "d2h_bytes_per_step": eb * 5,
                    "clips_per_step_per_gpu": eb, "host_numa": host_numa,
                    "per_rank_s_per_step": per_rank_e2e,
                    "h2d_gbs": e2e_val * 32000 / 1e9,
                    "h2d_ceiling_gbs": h2d_ceiling_gbs, "h2d_ceiling_per_rank_gbs": per_rank_h2d,
                    "h2d_ceiling_how": "every rank copies the same pinned buffer host->device at once, no kernels, "
                                       "max-over-ranks wall time",
                    "frac_of_h2d_ceiling": e2e_val * 32000 / 1e9 / h2d_ceiling_gbs,
                    "api": "WakeWordScorer.score_host -> ww_score_clips_host (pinned host "
                    "PCM in, host logits+decisions out)"},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "positives": int(dec.sum().item()), "keyword_hits": int(hits[:n_utt].sum().item()) if n_utt else 0,
            "fused_hbm_frac": (value / world) * FUSED_BYTES_PER_CLIP / 1e9 / measured_peaks()[0],
            "fused_path": {"algorithmic_bytes_per_clip": FUSED_BYTES_PER_CLIP, "dram_bytes_per_clip_ncu": 38758.6,
                           "ratio": 38758.6 / FUSED_BYTES_PER_CLIP,
                           "source": "profiles/r2_fused_dram_by_chunk.txt (ncu dram__bytes_read/write.sum over every kernel of "
                                     "3 fused passes, --cache-control none): the [chunk,13,63] fp32 features cross HBM once each way "
                                     "in the default (fastest) hand-over; `handoff` lists the two that keep them in L2"},
            "handoff": handoff,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY FAILURE in the measured run: %s\n" % json.dumps(parity))
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=1 << 20, help="clips per GPU resident in HBM")
    ap.add_argument("--e2e-clips", type=int, default=1 << 17, help="clips per e2e step (pinned host buffer)")
    ap.add_argument("--roofline-clips", type=int, default=1 << 20)
    ap.add_argument("--cpu-clips", type=int, default=8192, help="bounded CPU-baseline sample")
    ap.add_argument("--cnn", default="tensor", choices=["fp32", "tensor"],
                    help="tensor: tcgen05 fp16-operand CNN + exact fp32 re-score of borderline clips (default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-handoff", action="store_true", help="skip the comparison of the three feature hand-overs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the measured run")
    ap.add_argument("--parity-clips", type=int, default=PARITY_DISTINCT,
                    help="distinct clips per rank checked against the oracle (plus a 4096-clip strided subset)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
