"""GPU-box helper: the SAM ViT-H block's four GEMMs at the bench's chunk size (8 views -> M = 32768 / 39200 rows) with
the epilogues the model uses; run under `ncu --set full -k regex:gemm_bf16` for the roofline evidence, or plainly for
CUDA-event timings (prints TFLOP/s per shape)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from interactvlm_b200.ops import Context  # noqa: E402

ctx = Context(0)
E = 1280
M = 8 * 4096
Mw = 8 * 25 * 196
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).bfloat16()
x, xw = rnd(M, E), rnd(Mw, E)
wqkv, bqkv = rnd(3 * E, E, sc=E ** -0.5), rnd(3 * E)
wo, bo = rnd(E, E, sc=E ** -0.5), rnd(E)
w1, b1 = rnd(4 * E, E, sc=E ** -0.5), rnd(4 * E)
w2, b2 = rnd(E, 4 * E, sc=(4 * E) ** -0.5), rnd(E)
h = rnd(M, 4 * E)
cases = {
    "qkv_window": lambda: ctx.gemm(xw, wqkv, bias=bqkv, force_swap=-1),
    "proj": lambda: ctx.gemm(x, wo, bias=bo, residual=x, force_swap=-1),
    "mlp1_gelu": lambda: ctx.gemm(x, w1, bias=b1, act=1, force_swap=-1),
    "mlp2_res": lambda: ctx.gemm(h, w2, bias=b2, residual=x, force_swap=-1),
}
flops = {"qkv_window": 2 * Mw * 3 * E * E, "proj": 2 * M * E * E, "mlp1_gelu": 2 * M * 4 * E * E, "mlp2_res": 2 * M * 4 * E * E}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for name, fn in cases.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.4f} ms  {flops[name] / ms / 1e9:.1f} TFLOP/s", flush=True)

# ---- LLaMA prefill gate / up projection (8 x 329 rows): SwiGLU gate in the GEMM epilogue against GEMM + silu_mul
if len(sys.argv) > 2 and sys.argv[2] == "swiglu":
    from interactvlm_b200.layout import interleave_gate_up
    T, D, F = 2632, 5120, 13824
    xp = rnd(T, D)
    wil = interleave_gate_up(rnd(F, D, sc=D ** -0.5), rnd(F, D, sc=D ** -0.5))
    outp = torch.empty((T, F), device="cuda", dtype=torch.bfloat16)

    def t(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    a = t(lambda: ctx.gemm(xp, wil, act=5, out=outp))
    b = t(lambda: ctx.silu_mul(ctx.gemm(xp, wil), interleaved=True))
    c = t(lambda: ctx.gemm(xp, wil))
    fl = 2.0 * T * 2 * F * D
    print(f"prefill gate/up {T} x {2 * F} x {D}: fused SwiGLU epilogue {a * 1e3:.1f} us ({fl / a / 1e9:.0f} TFLOP/s); GEMM alone {c * 1e3:.1f} us "
          f"({fl / c / 1e9:.0f}); GEMM + silu_mul {b * 1e3:.1f} us", flush=True)

# ---- is the GELU epilogue what separates MLP-1 from the other shapes?  Same launch with each activation.
if len(sys.argv) > 2 and sys.argv[2] == "acts":
    for act, label in ((0, "none"), (3, "relu"), (2, "quick_gelu (1 MUFU pair)"), (1, "gelu (erf form)")):
        fn = lambda: ctx.gemm(x, w1, bias=b1, act=act, force_swap=-1)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"mlp1 M={M} N={4 * E} K={E} act={label}: {ms * 1e3:.1f} us  {flops['mlp1_gelu'] / ms / 1e9:.0f} TFLOP/s", flush=True)
