"""GPU-box helper: the decoder tail of the path at the bench batch (8 samples x 4 views): x4 bilinear upsample, lift from the
materialised 1024^2 logits, lift from the low-res logits (fused bilinear).  CUDA events; algorithmic bytes as SURVEY.md 8d
states them (B x 1.05 MB low-res logits + the CSR map + B x 27.6 KB out).  Under ncu, use -k regex:lift_warp|bilinear."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from interactvlm_b200 import synthetic as S  # noqa: E402
from interactvlm_b200.ops import LIFT_HUMAN, Context, LiftMap  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = Context(0)
p2v, bary = S.make_mesh_lift_maps(seed=0)
m = LiftMap(ctx, p2v, bary, S.N_SMPL)
g = torch.Generator().manual_seed(0)
low = (torch.randn(B, 4, 256, 256, generator=g) * 4).bfloat16().float().cuda()
full = ctx.bilinear(low.view(B * 4, 256, 256), 1024, 1024).view(B, 4, 1024, 1024)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, name, bytes_):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 (126 MB) between repetitions
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"{name:<46} {us:8.1f} us   {bytes_ / us / 1e3:8.1f} GB/s on {bytes_ / 1e6:.1f} MB algorithmic", flush=True)


csr = m.nnz * 8 + (4 * S.N_SMPL + 1) * 4
out_b = B * S.N_SMPL * 4
timed(lambda: ctx.bilinear(low.view(B * 4, 256, 256), 1024, 1024), "bilinear x4 (write 16.8 MB/sample)", B * 4 * (65536 + 1048576) * 4)
timed(lambda: m(full, LIFT_HUMAN), "lift from 1024^2 logits (warp per vertex-view)", csr + out_b + B * 4 * 1048576 * 4 * 0.18)
timed(lambda: m.lowres(low, LIFT_HUMAN), "lift from low-res logits (fused bilinear)", csr + out_b + B * 4 * 65536 * 4)
print(f"nnz {m.nnz}, CSR {csr / 1e6:.1f} MB, batch {B}")

# ---- SAM mask-decoder kernels at the bench shape (32 views): token->image attention (9 queries x 4096 keys, 8 heads of 16),
# image->token attention (4096 queries x 9 keys), second transposed conv + hypernetwork dot.  Variant 1 = the round-1 kernels.
nv, heads = B * 4, 8
rb = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).bfloat16().cuda()
qt, kt, vt = rb(nv, 9, 128), rb(nv, 4096, 128), rb(nv, 4096, 128)
up1, w2, b2, hyp = rb(nv, 4096, 4, 64), rb(4, 32, 64, sc=0.2), rb(32), rb(nv, 32)
for variant, label in ((0, "round 2"), (1, "round 1")):
    ctx.set_option("attn_small_variant", variant)
    timed(lambda: ctx.attn_small(qt, kt, vt, heads), f"attn 9 queries x 4096 keys ({label})", (2 * nv * 4096 * 128 + 2 * nv * 9 * 128) * 2)
    timed(lambda: ctx.upscale_hyper_dot(up1, w2, b2, hyp, nv, 64), f"upscale + hyper dot ({label})", nv * 4096 * 4 * 64 * 2 + nv * 65536 * 4)
ctx.set_option("attn_small_variant", 0)
timed(lambda: ctx.attn_small(kt, qt, qt, heads), "attn 4096 queries x 9 keys", 2 * nv * 4096 * 128 * 2)
