"""GPU-box helper: LayerNorm over SAM-sized activations ([rows, 1280] bf16), CUDA events, read + write bytes / time."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from interactvlm_b200.ops import Context  # noqa: E402

ctx = Context(0)
for rows, D in ((32768, 1280), (65536, 1280), (78400, 1280), (8 * 257, 1024)):
    xs = [torch.randn(rows, D, device="cuda").bfloat16() for _ in range(4)]   # rotate: 4 x 168 MB exceed L2
    g, b = torch.randn(D, device="cuda").bfloat16(), torch.randn(D, device="cuda").bfloat16()
    for i in range(3):
        ctx.layernorm(xs[i % 4], g, b, 1e-6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 40
    e0.record()
    for i in range(n):
        ctx.layernorm(xs[i % 4], g, b, 1e-6)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"layernorm [{rows}, {D}]: {us:7.2f} us  {2 * rows * D * 2 / us / 1e3:7.1f} GB/s (read + write)", flush=True)
