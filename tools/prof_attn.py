"""GPU-box helper: SAM attention shapes at the bench's chunk size (8 views): fused tcgen05 kernel vs the first-generation
path (rel-pos kernel + mma.sync flash attention).  Prints ms and TFLOP/s."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from interactvlm_b200.ops import Context  # noqa: E402

ctx = Context(0)
heads, hd = 16, 80
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).bfloat16()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
which = sys.argv[2] if len(sys.argv) > 2 else "both"


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, B, side in (("global_16views", 16, 64), ("window_16views", 400, 14)):
    S = side * side
    qkv = rnd(B * S, 3 * heads * hd, sc=0.5)
    rph, rpw = rnd(2 * side - 1, hd, sc=0.3), rnd(2 * side - 1, hd, sc=0.3)
    fl = 4.0 * B * heads * S * S * hd
    if which in ("both", "new"):
        ms = timeit(lambda: ctx.sam_attention(qkv, rph, rpw, B, heads, side, side, hd))
        print(f"{name} fused tcgen05: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        if side == 14:
            for ahead in (0, 148, 296, 592, 1184):
                ctx.set_option("attn_prefetch_ahead", ahead)
                ms = timeit(lambda: ctx.sam_attention(qkv, rph, rpw, B, heads, side, side, hd))
                print(f"{name} fused tcgen05, L2 prefetch {ahead} CTAs ahead: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            ctx.set_option("attn_prefetch_ahead", -1)
        if side == 14:
            for v_ in (3, 0, 3, 0):
                ctx.set_option("window_attn_variant", v_)
                ms = timeit(lambda: ctx.sam_attention(qkv, rph, rpw, B, heads, side, side, hd))
                print(f"{name} fused tcgen05, {'persistent CTAs' if v_ == 3 else 'one CTA per item'}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            ctx.set_option("window_attn_variant", 0)
        opt = "global_attn_variant" if side == 64 else "window_attn_variant"
        for variant, label in ((2, "one thread per row (round-1 kernel)" if side == 64 else "two threads per row"),
                               (1, "128-key tiles / 1 CTA per SM" if side == 64 else "tiled kernel")):
            ctx.set_option(opt, variant)
            ms = timeit(lambda: ctx.sam_attention(qkv, rph, rpw, B, heads, side, side, hd))
            ctx.set_option(opt, 0)
            print(f"{name} fused tcgen05, {label}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    if which in ("both", "old"):
        t = qkv.view(B, S, 3, heads, hd)

        def old():
            rel_h, rel_w = ctx.sam_relpos(qkv, rph, rpw, B, heads, side, side, hd)
            return ctx.attention(t[:, :, 0], t[:, :, 1], t[:, :, 2], hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=side, kw=side)

        ms = timeit(old)
        print(f"{name} relpos + mma.sync flash: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)

# ---- plain / causal attention of the path: LLaMA prefill (8 x 40 heads x 329 rows x 128, causal) and CLIP (8 x 16 x 257 x 64)
for name, B, H, S_, D, causal in (("llama_prefill_8x329", 8, 40, 329, 128, True), ("clip_8x257", 8, 16, 257, 64, False),
                                  ("llama_prefill_8x1024", 8, 40, 1024, 128, True)):
    q, k, v = (rnd(B, S_, H, D, sc=0.5) for _ in range(3))
    fl = 4.0 * B * H * S_ * S_ * D * (0.5 if causal else 1.0)
    for variant, label in ((0, "tcgen05"), (1, "mma.sync flash (round 1)")):
        ctx.set_option("attn_variant", variant)
        ms = timeit(lambda: ctx.attention(q, k, v, D ** -0.5, causal=causal))
        print(f"{name} {label}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    ctx.set_option("attn_variant", 0)
