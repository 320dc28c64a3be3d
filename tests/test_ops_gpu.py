"""Kernel-level numerics on a real B200: each sm_100a kernel against a plain PyTorch fp32 statement of the
same op on the same (bf16-representable) inputs.  Tolerances are written per test; integer/index work is exact.
All calls go through the C ABI (interactvlm_b200.ops -> ctypes -> libivlm_b200.so)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# bf16 output rounding alone is 2^-9 relative per element; fp32 accumulation order adds ~1e-6.
GEMM_TOL = 6e-3


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 128, 128), (256, 512, 1280), (1000, 3840, 1280), (777, 1280, 5120),
    (4900, 1280, 1280), (130, 32, 256), (300, 64, 64), (2640, 5120, 5120), (512, 8, 256),
])
def test_gemm_plain(ctx, M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    out = ctx.gemm(a, w, force_swap=-1)
    ref = a.float() @ w.float().t()
    assert out.shape == (M, N)
    assert rel_err(out, ref) < GEMM_TOL
    assert (out.float() - ref).abs().max().item() < 0.05 * ref.abs().max().item() + 1e-2


@pytest.mark.parametrize("M,N,K", [(8, 5120, 5120), (1, 256, 5120), (36, 2048, 256), (8, 32004, 512), (64, 15360, 5120)])
def test_gemm_swapped_small_m(ctx, M, N, K):
    a, w = rnd(M, K, seed=3), rnd(N, K, scale=K ** -0.5, seed=4)
    bias = rnd(N, seed=5)
    out = ctx.gemm(a, w, bias=bias, out_dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias.float()
    assert rel_err(out, ref) < 1e-4  # fp32 output: only accumulation-order noise


def test_gemm_epilogue_bias_act_residual(ctx):
    M, N, K = 1024, 5120, 1280
    a, w, bias, res = rnd(M, K, seed=6), rnd(N, K, scale=K ** -0.5, seed=7), rnd(N, seed=8), rnd(M, N, seed=9)
    for act, fn in [(1, torch.nn.functional.gelu), (2, lambda x: x * torch.sigmoid(1.702 * x)), (3, torch.relu)]:
        out = ctx.gemm(a, w, bias=bias, act=act, residual=res)
        y = (a.float() @ w.float().t() + bias.float()).bfloat16().float()
        ref = fn(y).bfloat16().float() + res.float()
        assert rel_err(out, ref) < GEMM_TOL, act


def test_gemm_row_map_and_res_mod(ctx):
    M, N, K = 512, 256, 128
    a, w = rnd(M, K, seed=10), rnd(N, K, scale=K ** -0.5, seed=11)
    perm = torch.randperm(M)
    row_map = perm.to(torch.int32)
    row_map[::7] = -1
    res = rnd(600, N, seed=12)
    out = torch.zeros(600, N, device=DEV, dtype=torch.bfloat16)
    ctx.gemm(a, w, residual=res, out=out, row_map=row_map.to(DEV))
    y = (a.float() @ w.float().t()).bfloat16().float()
    ref = torch.zeros(600, N, device=DEV)
    live = row_map >= 0
    ref[row_map[live].long().to(DEV)] = y[live.to(DEV)] + res.float()[row_map[live].long().to(DEV)]
    assert rel_err(out, ref) < GEMM_TOL
    # broadcast residual (pos_embed style)
    tab = rnd(128, N, seed=13)
    out2 = ctx.gemm(a, w, residual=tab, res_row_mod=128)
    ref2 = y + tab.float().repeat(M // 128, 1)
    assert rel_err(out2, ref2) < GEMM_TOL


def test_gemm_split_k(ctx):
    M, N, K = 8, 5120, 13824
    a, w = rnd(M, K, seed=14), rnd(N, K, scale=K ** -0.5, seed=15)
    acc = ctx.gemm(a, w, out_dtype=torch.float32, k_splits=8)
    ref = a.float() @ w.float().t()
    assert rel_err(acc, ref) < 1e-4
    res = rnd(M, N, seed=16)
    out = ctx.finalize(acc, residual=res)
    assert rel_err(out, ref.bfloat16().float() + res.float()) < GEMM_TOL


def test_gemm_k_tail_and_strided(ctx):
    # K not a multiple of 64 (TMA zero fill) and a strided activation view (packed qkv column slice)
    M, N, K = 257, 136, 88
    big = rnd(M, 3 * K, seed=17)
    a = big[:, K:2 * K]
    w = rnd(N, K, seed=18)
    out = ctx.gemm(a, w)
    assert rel_err(out, a.float() @ w.float().t()) < GEMM_TOL


def test_layernorm_rmsnorm(ctx):
    for D in (256, 1024, 1280):
        x, g, b = rnd(999, D, seed=19), rnd(D, seed=20), rnd(D, seed=21)
        y = ctx.layernorm(x, g, b, 1e-6)
        ref = torch.nn.functional.layer_norm(x.float(), (D,), g.float(), b.float(), 1e-6)
        assert rel_err(y, ref) < 4e-3
    x = rnd(50, 64, seed=22)
    g, b = rnd(64, seed=23), rnd(64, seed=24)
    y = ctx.layernorm(x, g, b, 1e-6, act=1)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x.float(), (64,), g.float(), b.float(), 1e-6).bfloat16().float())
    assert rel_err(y, ref) < 4e-3
    # gather + zero pad rows
    rm = torch.tensor([3, -1, 0, 7, -1], dtype=torch.int32, device=DEV)
    x = rnd(8, 256, seed=25); g, b = rnd(256, seed=26), rnd(256, seed=27)
    y = ctx.layernorm(x, g, b, 1e-6, row_map=rm)
    ref = torch.nn.functional.layer_norm(x.float(), (256,), g.float(), b.float(), 1e-6)
    assert torch.equal(y[1], torch.zeros_like(y[1])) and torch.equal(y[4], torch.zeros_like(y[4]))
    assert rel_err(y[[0, 2, 3]], ref[[3, 0, 7]]) < 4e-3
    x, g = rnd(333, 5120, seed=28), rnd(5120, seed=29)
    y = ctx.rmsnorm(x, g, 1e-5)
    xf = x.float()
    ref = g.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float()
    assert rel_err(y, ref) < 4e-3


def test_elementwise(ctx):
    a, b = rnd(4, 64, 256, seed=30), rnd(64, 256, seed=31)
    assert torch.equal(ctx.add_bcast(a, b), (a.float() + b.float()).bfloat16())
    gu = rnd(37, 2 * 1024, seed=32)
    ref = (torch.nn.functional.silu(gu[:, :1024].float()).bfloat16().float() * gu[:, 1024:].float())
    assert rel_err(ctx.silu_mul(gu), ref) < 4e-3
    x = torch.randn(1000, device=DEV)
    assert torch.equal(ctx.to_bf16(x), x.bfloat16())
    assert torch.equal(ctx.to_f32(x.bfloat16()), x.bfloat16().float())


def test_im2col(ctx):
    img = rnd(2, 3, 28, 42, seed=33)
    cols = ctx.im2col_patch(img, 14, ldk=592)
    ref = torch.nn.functional.unfold(img.float(), 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(cols[:, :588].float(), ref) and cols[:, 588:].abs().max().item() == 0
    # patch sizes that are multiples of 8 take the 16-byte-per-thread kernel (SAM: 16 x 16 patches), with and without a padded pitch
    img = rnd(3, 3, 64, 96, seed=63)
    ref = torch.nn.functional.unfold(img.float(), 16, stride=16).transpose(1, 2).reshape(-1, 768)
    assert torch.equal(ctx.im2col_patch(img, 16).float(), ref)
    cols = ctx.im2col_patch(img, 16, ldk=776)
    assert torch.equal(cols[:, :768].float(), ref) and cols[:, 768:].abs().max().item() == 0
    img = rnd(2, 3, 48, 24, seed=64)
    ref = torch.nn.functional.unfold(img.float(), 8, stride=8).transpose(1, 2).reshape(-1, 192)
    assert torch.equal(ctx.im2col_patch(img, 8).float(), ref)
    x = rnd(2, 6, 5, 16, seed=34)  # [N,H,W,C]
    cols = ctx.im2col_3x3(x.contiguous(), 2, 6, 5)
    ref = torch.nn.functional.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1)  # [N, C*9, HW] with k = c*9 + tap
    ref = ref.view(2, 16, 9, 30).permute(0, 3, 2, 1).reshape(60, 9 * 16)  # -> k = tap*C + c
    assert torch.equal(cols.float(), ref)


@pytest.mark.parametrize("B,H,Sq,Sk,D,causal", [
    (2, 16, 257, 257, 64, False), (3, 4, 330, 330, 128, True), (2, 2, 196, 196, 80, False),
    (1, 3, 64, 200, 128, True), (1, 2, 1000, 1000, 80, False), (2, 2, 33, 33, 64, True),
])
def test_flash_attention(ctx, B, H, Sq, Sk, D, causal):
    qkv = rnd(B, max(Sq, Sk), 3, H, D, seed=35)
    q, k, v = qkv[:, :Sq, 0], qkv[:, :Sk, 1], qkv[:, :Sk, 2]
    scale = D ** -0.5
    out = ctx.attention(q, k, v, scale, causal=causal)
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * scale
    if causal:
        mask = torch.ones(Sq, Sk, device=DEV).tril(Sk - Sq).bool()
        s = s.masked_fill(~mask, float("-inf"))
    ref = (s.softmax(-1) @ vf).permute(0, 2, 1, 3)
    assert rel_err(out, ref) < 8e-3
    assert (out.float() - ref).abs().max().item() < 3e-2
    if D in (64, 128) and Sq == Sk:
        # this layout runs on flash_attn_tcgen05_kernel; option attn_variant = 1 selects the mma.sync kernel it replaces
        ctx.set_option("attn_variant", 1)
        old = ctx.attention(q, k, v, scale, causal=causal)
        ctx.set_option("attn_variant", 0)
        assert rel_err(old, ref) < 8e-3 and rel_err(out, old) < 8e-3


@pytest.mark.parametrize("B,H,S,D,causal", [(8, 40, 329, 128, True), (8, 16, 257, 64, False), (1, 2, 128, 128, True), (3, 2, 129, 64, True),
                                            (2, 3, 700, 128, True), (2, 2, 640, 64, False), (5, 1, 1, 128, True)])
def test_flash_attention_tcgen05(ctx, B, H, S, D, causal):
    """flash_attn_tcgen05_kernel at the path's shapes (LLaMA prefill: 8 x 40 heads x 329 rows, causal; CLIP: 257 tokens) and at
    tile edges (S = 128, 129, several key tiles, one token), separate q / k / v matrices with their own pitches (the prefill
    layout) and a packed qkv matrix (the CLIP layout), against fp32 torch."""
    C = H * D
    for packed in (False, True):
        if packed:
            qkv = rnd(B, S, 3, H, D, seed=70)
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        else:
            q, k, v = (rnd(B, S, H, D, seed=71 + i) for i in range(3))
        out = ctx.attention(q, k, v, D ** -0.5, causal=causal)
        qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
        sc = qf @ kf.transpose(-1, -2) * D ** -0.5
        if causal:
            sc = sc.masked_fill(~torch.ones(S, S, device=DEV).tril().bool(), float("-inf"))
        ref = (sc.softmax(-1) @ vf).permute(0, 2, 1, 3)
        assert torch.isfinite(out.float()).all()
        assert rel_err(out, ref) < 8e-3, (packed, rel_err(out, ref))
        assert (out.float() - ref).abs().max().item() < 3e-2
        n0 = ctx.launch_count()
        ctx.set_option("attn_variant", 1)
        old = ctx.attention(q, k, v, D ** -0.5, causal=causal)
        ctx.set_option("attn_variant", 0)
        assert rel_err(out, old) < 8e-3


@pytest.mark.parametrize("Hq,Wq,B", [(14, 14, 5), (64, 64, 1)])
def test_sam_relpos_attention(ctx, Hq, Wq, B):
    heads, hd = 2, 80
    S = Hq * Wq
    qkv = rnd(B * S, 3 * heads * hd, seed=36, scale=0.5)
    rph, rpw = rnd(2 * Hq - 1, hd, seed=37, scale=0.3), rnd(2 * Wq - 1, hd, seed=38, scale=0.3)
    rel_h, rel_w = ctx.sam_relpos(qkv, rph, rpw, B, heads, Hq, Wq, hd)
    t = qkv.view(B, S, 3, heads, hd)
    q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    out = ctx.attention(q, k, v, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=Hq, kw=Wq)
    # reference: image_encoder.py:235-260 / :354-392 in fp32
    qf = q.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    kf = k.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    vf = v.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    idx_h = (torch.arange(Hq)[:, None] - torch.arange(Hq)[None, :] + Hq - 1).to(DEV)
    idx_w = (torch.arange(Wq)[:, None] - torch.arange(Wq)[None, :] + Wq - 1).to(DEV)
    Rh, Rw = rph.float()[idx_h], rpw.float()[idx_w]
    rq = qf.view(B * heads, Hq, Wq, hd)
    rh = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rw = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    assert rel_err(rel_h.view(B * heads, Hq, Wq, Hq), rh) < 4e-3
    assert rel_err(rel_w.view(B * heads, Hq, Wq, Wq), rw) < 4e-3
    attn = (qf * hd ** -0.5) @ kf.transpose(-1, -2)
    attn = (attn.view(B * heads, Hq, Wq, Hq, Wq) + rh[..., None] + rw[..., None, :]).view(B * heads, S, S)
    ref = (attn.softmax(-1) @ vf).view(B, heads, S, hd).permute(0, 2, 1, 3)
    assert rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("Hq,Wq,B,heads", [(64, 64, 2, 2), (14, 14, 5, 2), (14, 14, 50, 16)])
def test_sam_attention_tcgen05(ctx, Hq, Wq, B, heads):
    """Fused tcgen05 attention with in-kernel rel-pos tables vs the fp32 statement of image_encoder.py:235-260/:354-392."""
    hd = 80
    S = Hq * Wq
    qkv = rnd(B * S, 3 * heads * hd, seed=60, scale=0.5)
    rph, rpw = rnd(2 * Hq - 1, hd, seed=61, scale=0.3), rnd(2 * Wq - 1, hd, seed=62, scale=0.3)
    out = ctx.sam_attention(qkv, rph, rpw, B, heads, Hq, Wq, hd)
    torch.cuda.synchronize()
    t = qkv.view(B, S, 3, heads, hd)
    q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    qf = q.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    kf = k.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    vf = v.float().permute(0, 2, 1, 3).reshape(B * heads, S, hd)
    idx_h = (torch.arange(Hq)[:, None] - torch.arange(Hq)[None, :] + Hq - 1).to(DEV)
    idx_w = (torch.arange(Wq)[:, None] - torch.arange(Wq)[None, :] + Wq - 1).to(DEV)
    rq = qf.view(B * heads, Hq, Wq, hd)
    rh = torch.einsum("bhwc,hkc->bhwk", rq, rph.float()[idx_h]).bfloat16().float()
    rw = torch.einsum("bhwc,wkc->bhwk", rq, rpw.float()[idx_w]).bfloat16().float()
    attn = (qf * hd ** -0.5) @ kf.transpose(-1, -2)
    attn = (attn.view(B * heads, Hq, Wq, Hq, Wq) + rh[..., None] + rw[..., None, :]).view(B * heads, S, S)
    ref = (attn.softmax(-1) @ vf).view(B, heads, S, hd).permute(0, 2, 1, 3).reshape(B * S, heads * hd)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    assert (out.float() - ref).abs().max().item() < 3e-2
    # the other kernel variants must agree with the default (two threads per query row): 64-key tiles with one thread per
    # row (2), 128-key tiles / one CTA per SM (global, 1); one thread per row (window, 2), tiled kernel (window, 1)
    opt = "global_attn_variant" if Hq == 64 else "window_attn_variant"
    for variant in (1, 2):
        ctx.set_option(opt, variant)
        alt = ctx.sam_attention(qkv, rph, rpw, B, heads, Hq, Wq, hd)
        ctx.set_option(opt, 0)
        assert rel_err(alt, ref) < 1e-2 and rel_err(alt, out) < 5e-3
    if Hq == 14:  # window_unpartition fused into the store: rows go to their token positions, padding rows are dropped
        gmap = torch.randperm(B * S + 40, generator=torch.Generator().manual_seed(3))[: B * S].to(torch.int32)
        gmap[::7] = -1
        gmap = gmap.to(DEV)
        scat = ctx.sam_attention(qkv, rph, rpw, B, heads, Hq, Wq, hd, out_map=gmap, out_rows=B * S + 40,
                                 out=torch.zeros(B * S + 40, heads * hd, device=DEV, dtype=torch.bfloat16))
        live = gmap >= 0
        assert torch.equal(scat[gmap[live].long()], out[live])
        dead = torch.ones(B * S + 40, dtype=torch.bool, device=DEV)
        dead[gmap[live].long()] = False
        assert scat[dead].abs().max().item() == 0
        # the persistent form of the window kernel (variant 3: CTAs loop over the items, next loads under the previous epilogue)
        # and the one-CTA-per-item form run the same arithmetic in the same order
        for other in ([3] if torch.equal(out, out) else []):
            base_variant = 0
            for v_ in (other, base_variant):
                ctx.set_option("window_attn_variant", v_)
                res = ctx.sam_attention(qkv, rph, rpw, B, heads, Hq, Wq, hd)
                res_sc = ctx.sam_attention(qkv, rph, rpw, B, heads, Hq, Wq, hd, out_map=gmap, out_rows=B * S + 40,
                                           out=torch.zeros(B * S + 40, heads * hd, device=DEV, dtype=torch.bfloat16))
                if v_ == other:
                    pers, pers_sc = res, res_sc
            ctx.set_option("window_attn_variant", 0)
            assert torch.equal(pers, res) and torch.equal(pers_sc, res_sc)
    # and against the first-generation path (separate rel-pos kernel + mma.sync flash attention)
    rel_h, rel_w = ctx.sam_relpos(qkv, rph, rpw, B, heads, Hq, Wq, hd)
    old = ctx.attention(q, k, v, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=Hq, kw=Wq).reshape(B * S, heads * hd)
    assert rel_err(out, old) < 1e-2


def test_attn_small(ctx):
    heads = 8
    for (Bq, B, Nq, Nk, Cc) in [(1, 4, 9, 4096, 128), (4, 4, 9, 9, 256), (4, 4, 4096, 9, 128), (2, 2, 9, 4096, 128)]:
        q, k, v = rnd(Bq, Nq, Cc, seed=39), rnd(B, Nk, Cc, seed=40), rnd(B, Nk, Cc, seed=41)
        out = ctx.attn_small(q, k, v, heads)
        hd = Cc // heads
        qf = q.float().expand(B, Nq, Cc).reshape(B, Nq, heads, hd).transpose(1, 2)
        kf = k.float().reshape(B, Nk, heads, hd).transpose(1, 2)
        vf = v.float().reshape(B, Nk, heads, hd).transpose(1, 2)
        ref = ((qf @ kf.transpose(-1, -2) / math.sqrt(hd)).softmax(-1) @ vf).transpose(1, 2).reshape(B, Nq, Cc)
        assert rel_err(out, ref) < 2e-2, (Bq, B, Nq, Nk, Cc)
        if Nk > 16 and hd == 16:
            # all-queries-in-one-sweep kernel against the per-query form it replaces: same rounding points and reduction order
            ctx.set_option("attn_small_variant", 1)
            old = ctx.attn_small(q, k, v, heads)
            ctx.set_option("attn_small_variant", 0)
            assert (out.float() - old.float()).abs().max().item() <= 2 ** -8 * old.float().abs().max().item()
            assert rel_err(out, old) < 1e-3
    # more than 10 queries (second instantiation), ragged key count
    q, k, v = rnd(3, 13, 128, seed=58), rnd(3, 1000, 128, seed=59), rnd(3, 1000, 128, seed=60)
    out = ctx.attn_small(q, k, v, heads)
    ctx.set_option("attn_small_variant", 1)
    old = ctx.attn_small(q, k, v, heads)
    ctx.set_option("attn_small_variant", 0)
    assert rel_err(out, old) < 1e-3


def test_llm_glue(ctx):
    vocab, D, B, Lq, n_img = 100, 64, 3, 12, 5
    embed, img = rnd(vocab, D, seed=42), rnd(B, n_img, D, seed=43)
    ids = torch.randint(0, vocab, (B, Lq), dtype=torch.int32)
    pos = [4, 0, 11]
    for b, p_ in enumerate(pos):
        ids[b, p_] = -200
    out = ctx.embed_splice(embed, ids.to(DEV), img)
    for b, p_ in enumerate(pos):
        ref = torch.cat([embed[ids[b, :p_].long()], img[b], embed[ids[b, p_ + 1:].long()]], 0)
        assert torch.equal(out[b], ref)
    idx = torch.tensor([5, 0, 99], dtype=torch.int32, device=DEV)
    assert torch.equal(ctx.embed_gather(embed, idx), embed[idx.long()])
    assert torch.equal(ctx.gather_rows(embed, idx), embed[idx.long()])
    logits = torch.randn(5, 32004, device=DEV)
    logits[2, 77] = logits[2, 900] = 100.0
    am = ctx.argmax(logits)
    assert torch.equal(am.long(), logits.argmax(-1)) and am[2].item() == 77


def _rope_ref(x, cos, sin):  # HF 4.31 apply_rotary_pos_emb in bf16
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    rot = torch.cat((-x2, x1), -1)
    return (x * cos) + (rot * sin)


def test_rope_paged_decode(ctx):
    H, hd, page, T = 4, 128, 16, 37
    D = H * hd
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    fr = torch.outer(torch.arange(64).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos_t, sin_t = emb.cos().bfloat16().to(DEV), emb.sin().bfloat16().to(DEV)
    qkv = rnd(T, 3 * D, seed=44)
    positions = torch.arange(T, dtype=torch.int32, device=DEV)
    pages = torch.tensor([2, 0, 1], dtype=torch.int32)
    slot = (pages[(positions.cpu() // page).long()] * page + positions.cpu() % page).to(torch.int32).to(DEV)
    kc = torch.zeros(3, H, page, hd, device=DEV, dtype=torch.bfloat16)  # [pages, H, page, hd]
    vc = torch.zeros_like(kc)
    q_out, k_out, v_out = ctx.rope_kv_store(qkv, positions, slot, cos_t, sin_t, H, hd, kc, vc, page_size=page)
    q, k, v = (qkv[:, i * D:(i + 1) * D].view(T, H, hd) for i in range(3))
    c, s = cos_t[:T, None, :], sin_t[:T, None, :]
    assert torch.equal(q_out.view(T, H, hd), _rope_ref(q, c, s))
    assert torch.equal(k_out.view(T, H, hd), _rope_ref(k, c, s))
    assert torch.equal(v_out.view(T, H, hd), v)
    sl = slot.long()
    assert torch.equal(kc[sl // page, :, sl % page], k_out.view(T, H, hd)) and torch.equal(vc[sl // page, :, sl % page], v)
    # decode attention for a "next token" query over those T keys
    qd = rnd(1, D, seed=45)
    bt = pages.view(1, 3).to(DEV)
    out = ctx.decode_attention(qd, kc, vc, bt, torch.tensor([T], dtype=torch.int32, device=DEV), H, hd, page)
    kf, vf = k_out.view(T, H, hd).float(), v.float()
    sc = torch.einsum("hd,thd->ht", qd.view(H, hd).float(), kf) / math.sqrt(hd)
    ref = torch.einsum("ht,thd->hd", sc.softmax(-1), vf).reshape(1, D)
    assert rel_err(out, ref) < 1e-2


def test_bilinear_matches_torch(ctx):
    src = torch.randn(3, 256, 256, device=DEV)
    out = ctx.bilinear(src, 1024, 1024)
    ref = torch.nn.functional.interpolate(src[None], (1024, 1024), mode="bilinear", align_corners=False)[0]
    assert (out - ref).abs().max().item() < 1e-5
    out = ctx.bilinear(src, 300, 500, crop_h=200, crop_w=256)
    ref = torch.nn.functional.interpolate(src[None, :, :200, :256], (300, 500), mode="bilinear", align_corners=False)[0]
    assert (out - ref).abs().max().item() < 1e-5
    # the row-blocked float4 kernel (aligned output, width % 4 == 0) and the generic one (here: a 4-byte-offset output) agree bit for bit
    buf = torch.empty(3 * 300 * 500 + 1, device=DEV)
    generic = ctx.bilinear(src, 300, 500, crop_h=200, crop_w=256, out=buf[1:].view(3, 300, 500))
    assert torch.equal(generic, out)
    out = ctx.bilinear(src, 77, 133)                                            # width % 4 != 0 -> generic kernel
    ref = torch.nn.functional.interpolate(src[None], (77, 133), mode="bilinear", align_corners=False)[0]
    assert (out - ref).abs().max().item() < 1e-5


def test_cam_gate(ctx):
    B, V = 3, 4
    cam, emb = rnd(B, V, 5, seed=46), rnd(B, 256, seed=47)
    w1, b1, w2, b2 = rnd(128, 5, seed=48), rnd(128, seed=49), rnd(128, 128, scale=0.1, seed=50), rnd(128, seed=51)
    wv, bv = rnd(V, 256, 128, scale=0.1, seed=52), rnd(V, 256, seed=53)
    out = ctx.cam_gate(cam, emb, w1, b1, w2, b2, wv, bv)
    h = torch.relu(cam.float() @ w1.float().t() + b1.float()).bfloat16().float()
    h = torch.relu(h @ w2.float().t() + b2.float()).bfloat16().float()
    g = torch.sigmoid((torch.einsum("bvk,vnk->bvn", h, wv.float()) + bv.float()).bfloat16().float())
    ref = emb.float()[:, None, :] * g
    assert rel_err(out, ref) < 1e-2


def test_upscale_hyper_dot(ctx):
    Bv, G = 2, 8
    up1 = rnd(Bv, G * G, 4, 64, seed=54)
    w2, b2, hyper = rnd(4, 32, 64, scale=0.2, seed=55), rnd(32, seed=56), rnd(Bv, 32, seed=57)
    out = ctx.upscale_hyper_dot(up1, w2, b2, hyper, Bv, G)
    z = torch.einsum("btpk,qck->btpqc", up1.float(), w2.float()) + b2.float()
    z = torch.nn.functional.gelu(z.bfloat16().float()).bfloat16().float()
    m = torch.einsum("btpqc,bc->btpq", z, hyper.float())  # [Bv, G*G, p1, p2]
    m = m.view(Bv, G, G, 2, 2, 2, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(Bv, G * 4, G * 4)
    assert rel_err(out, m) < 1e-2
    ctx.set_option("attn_small_variant", 1)                 # the scalar form of the kernel (A/B option)
    old = ctx.upscale_hyper_dot(up1, w2, b2, hyper, Bv, G)
    ctx.set_option("attn_small_variant", 0)
    assert rel_err(out, old) < 5e-3
    for Bv, G in ((3, 64), (1, 3)):                         # full SAM grid; a grid whose pixel count is not a multiple of 16
        up1 = rnd(Bv, G * G, 4, 64, seed=61)
        hyper = rnd(Bv, 32, seed=62)
        out = ctx.upscale_hyper_dot(up1, w2, b2, hyper, Bv, G)
        z = torch.einsum("btpk,qck->btpqc", up1.float(), w2.float()) + b2.float()
        z = torch.nn.functional.gelu(z.bfloat16().float()).bfloat16().float()
        m = torch.einsum("btpqc,bc->btpq", z, hyper.float())
        m = m.view(Bv, G, G, 2, 2, 2, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(Bv, G * 4, G * 4)
        assert rel_err(out, m) < 1e-2, (Bv, G)


@pytest.mark.parametrize("M,F,K", [(2632, 13824, 5120), (130, 704, 256), (300, 40, 512), (65, 1416, 64)])
def test_gemm_swiglu_epilogue(ctx, M, F, K):
    """ivlm_gemm_bf16 with IVLM_ACT_SWIGLU (interleaved gate / up rows, the SwiGLU gate applied in the staged epilogue: [M, F]
    written instead of [M, 2F]) against the two launches it replaces and against fp32 torch."""
    from interactvlm_b200.layout import interleave_gate_up
    x = rnd(M, K, seed=80)
    gate, up = rnd(F, K, scale=K ** -0.5, seed=81), rnd(F, K, scale=K ** -0.5, seed=82)
    wil = interleave_gate_up(gate, up)
    out = ctx.gemm(x, wil, act=5, out=torch.empty((M, F), device=DEV, dtype=torch.bfloat16))
    two = ctx.silu_mul(ctx.gemm(x, wil), interleaved=True)
    g, u = (x.float() @ gate.float().t()).bfloat16().float(), (x.float() @ up.float().t()).bfloat16().float()
    ref = torch.nn.functional.silu(g).bfloat16().float() * u
    assert out.shape == (M, F) and torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 4e-3 and rel_err(out, two) < 2e-3
    # the SFU silu of the epilogue against the expf silu of silu_mul_kernel: at most one bf16 ulp of the product apart
    assert ((out.float() - two.float()).abs() <= 2 ** -7 * two.float().abs() + 1e-6).all()
