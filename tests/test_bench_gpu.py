"""bench.py's product arm on the GPU at the tiny configuration with its DEFAULT flags (per-kernel profiling pass on): one
process, and two ranks under torchrun (skipped with fewer than two devices) -- the multi-rank path must not issue a
collective from a subset of the ranks (round-1 defect: the rank-0-only profiling step all-gathered while the others sat in
the final barrier)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _line(stdout):
    rows = [l for l in stdout.strip().splitlines() if l.startswith("{")]
    assert rows, stdout[-2000:]
    return json.loads(rows[-1])


def test_bench_tiny_one_gpu_default_flags():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "tiny", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 100
    assert d["roofline"]["achieved"] > 0 and d["cpu_baseline"]["kind"] == "port" and "not extrapolated" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0


@pytest.mark.parametrize("workload", ["hcontact", "oafford_pc"])
def test_bench_tiny_two_ranks_default_flags(workload):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(ROOT / "bench.py"), "--gpus", "2", "--config", "tiny", "--steps", "2", "--warmup", "3",
                        "--workload", workload],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT), env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    d = _line(r.stdout)
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["config"]["global_batch"] == 16
    assert "roofline" in d and d["roofline"]["achieved"] > 0


def test_bench_tiny_oafford_and_sweep_one_gpu():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "tiny", "--steps", "2", "--warmup", "3", "--workload", "oafford_pc",
                        "--sweep", "1,4", "--no-profile"], capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    rows = [json.loads(l) for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert [x["config"]["batch_per_gpu"] for x in rows] == [1, 4] and all(x["value"] > 0 for x in rows)
    assert all("oafford_pc" in x["metric"] for x in rows)


def test_bench_tiny_joint_fit_one_gpu():
    env = dict(os.environ, IVLM_FIT_ITERS="12")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "tiny", "--steps", "1", "--warmup", "1", "--batch", "2",
                        "--workload", "joint_fit"], capture_output=True, text=True, timeout=900, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    assert d["unit"] == "samples/s" and d["value"] > 0 and d["e2e"]["value"] > 0 and d["config"]["fit_iterations"] == 12
    assert d["gpu_launches"] > 200 and d["fit_ms_per_iteration"] > 0
