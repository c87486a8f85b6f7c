"""bench.py host-side pieces that must behave without a GPU: the reference arm's JSON line, the clock sampler and the
per-rank host placement (both best effort: they report what they could not do instead of failing)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    import bench

    return bench


def test_clock_sampler_reports_zero_samples_without_a_gpu():
    bench = _bench()
    c = bench.ClockSampler(0)
    c.start()
    time.sleep(0.05)
    got = c.stop(time.time() - 0.05, time.time())
    assert set(got) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    assert got["samples"] == 0 and got["sm_mhz"] is None and got["reasons"] == []


def test_host_placement_is_best_effort():
    bench = _bench()
    before = os.sched_getaffinity(0)
    info = bench.bind_host_to_gpu_node(0)
    assert info["mempolicy"] == "unchanged" and info["node"] is None and "error" in info
    assert os.sched_getaffinity(0) == before            # nothing was bound


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` needs no GPU: one JSON line with the keys the driver reads (tiny sample here)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-clips", "400"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_under_torchrun_prints_on_rank0_only():
    """N > 1: the driver launches the reference arm with torchrun; rank 0 alone works and prints, the others exit 0."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-clips", "400"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["cores"] >= 1          # all host threads, not torchrun's OMP_NUM_THREADS=1
