"""bench.py's reference arm (the CPU restatement of the reference algorithm timed on the host cores) runs without a GPU and
prints the contract's JSON line; here at the tiny configuration so that it takes seconds.  The product arm must refuse to
run without CUDA (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"].startswith("images/sec") and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "WITHOUT KV cache" in cb["sample"] and "not extrapolated" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    import os

    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "tiny", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
