"""ivlm_decode_linear (csrc/decode_stream.cu) on a B200: the fused decode-step layers against the unfused chain of kernels they
replace and against fp32 torch restatements of the HF ops (LlamaRMSNorm, Linear, apply_rotary_pos_emb, SwiGLU), at shapes that
exercise every decomposition case -- ragged last tile (N % 16 != 0), ragged last stage (K % 512 != 0, K % 64 == 0), fewer tiles than SMs,
ranges cut inside tiles (partial-tile hand-over), streamed and resident activations, M < 8 -- plus determinism across launches."""
import math

import pytest
import torch

from interactvlm_b200.layout import interleave_gate_up, pair_rows

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16().to(DEV)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def rms_ref(x, g, eps):
    xf = x.float()
    return (g.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float()).bfloat16()


@pytest.mark.parametrize("M,N,K", [(8, 5120, 5120), (8, 5120, 13824), (5, 2000, 1088), (1, 48, 256), (8, 32004, 5120), (3, 4096, 576),
                                   (8, 320, 8256)])
def test_plain_epilogue_vs_torch_and_gemm(ctx, M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    res, bias = rnd(M, N, seed=3), rnd(N, seed=4)
    ref = (a.float() @ w.float().t() + bias.float()).bfloat16().float() + res.float()
    out = ctx.decode_linear(a, w, bias=bias, residual=res)
    assert rel_err(out, ref) < 4e-3
    base = ctx.gemm(a, w, bias=bias, residual=res)          # the kernels it replaces (other fp32 summation order)
    assert rel_err(out, base) < 4e-3 and (out.float() - base.float()).abs().max().item() <= 0.07 * ref.abs().max().item()
    for _ in range(3):                                      # deterministic, flags self-reset between launches
        assert torch.equal(ctx.decode_linear(a, w, bias=bias, residual=res), out)
    logits = ctx.decode_linear(a, w, out_dtype=torch.float32)
    assert rel_err(logits, a.float() @ w.float().t()) < 2e-3


def test_fused_rmsnorm_prologue(ctx):
    M, N, K = 8, 1024, 5120
    x, g, w = rnd(M, K, seed=5), (1 + 0.1 * torch.randn(K, generator=torch.Generator().manual_seed(6))).bfloat16().to(DEV), rnd(N, K, scale=K ** -0.5, seed=7)
    y = ctx.rmsnorm(x, g, 1e-5)
    assert torch.equal(y, rms_ref(x, g, 1e-5))
    out = ctx.decode_linear(x, w, gamma=g, eps=1e-5)
    assert torch.equal(out, ctx.decode_linear(y, w))         # the prologue produces exactly the rmsnorm kernel's rows
    assert rel_err(out, y.float() @ w.float().t()) < 4e-3


def test_swiglu_epilogue(ctx):
    M, F, K = 7, 1376, 512
    x = rnd(M, K, seed=8)
    gate, up = rnd(F, K, scale=K ** -0.5, seed=9), rnd(F, K, scale=K ** -0.5, seed=10)
    wil = interleave_gate_up(gate, up)
    out = ctx.decode_linear(x, wil, epilogue=1)
    gu = ctx.decode_linear(x, torch.cat([gate, up], 0))      # same accumulation order per row -> bit-comparable
    ref = (torch.nn.functional.silu(gu[:, :F].float()).bfloat16().float() * gu[:, F:].float()).bfloat16()
    assert out.shape == (M, F) and rel_err(out, ref) < 2e-3
    assert torch.equal(out, ctx.silu_mul(ctx.decode_linear(x, wil), interleaved=True))
    assert torch.equal(ctx.silu_mul(ctx.gemm(x, wil), interleaved=True), ctx.silu_mul(ctx.gemm(x, torch.cat([gate, up], 0))))


def test_rope_kv_epilogue_matches_the_separate_kernels(ctx):
    H, hd, page, M, K = 4, 128, 16, 6, 512
    D = H * hd
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    fr = torch.outer(torch.arange(128).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos_t, sin_t = emb.cos().bfloat16().to(DEV), emb.sin().bfloat16().to(DEV)
    x, gam = rnd(M, K, seed=11), (1 + 0.1 * torch.randn(K, generator=torch.Generator().manual_seed(12))).bfloat16().to(DEV)
    wq, wk, wv = (rnd(D, K, scale=K ** -0.5, seed=s) for s in (13, 14, 15))
    w_nat = torch.cat([wq, wk, wv], 0)
    w_pair = torch.cat([pair_rows(wq, H, hd), pair_rows(wk, H, hd), wv], 0)
    positions = torch.tensor([5, 17, 33, 2, 70, 71], dtype=torch.int32, device=DEV)
    pages_per = 8
    slot = (torch.arange(M, dtype=torch.int32, device=DEV) * pages_per * page + positions).contiguous()
    kc = torch.zeros(M * pages_per, H, page, hd, device=DEV, dtype=torch.bfloat16)
    vc, kc2, vc2 = torch.zeros_like(kc), torch.zeros_like(kc), torch.zeros_like(kc)
    rope = dict(positions=positions, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc, v_cache=vc, H=H, hd=hd, page_size=page)
    q = ctx.decode_linear(x, w_pair, gamma=gam, eps=1e-5, epilogue=2, rope=rope)
    # the chain it replaces, on the natural layout, with the same GEMM kernel (identical accumulation order per row)
    qkv = ctx.decode_linear(ctx.rmsnorm(x, gam, 1e-5), w_nat)
    q2, _, _ = ctx.rope_kv_store(qkv, positions, slot, cos_t, sin_t, H, hd, kc2, vc2, want_kv=False, page_size=page)
    assert torch.equal(q, q2) and torch.equal(kc, kc2) and torch.equal(vc, vc2)
    # and the paired flag of the stand-alone kernel (prefill path)
    kc3, vc3 = torch.zeros_like(kc), torch.zeros_like(kc)
    q3, k3, v3 = ctx.rope_kv_store(ctx.decode_linear(ctx.rmsnorm(x, gam, 1e-5), w_pair), positions, slot, cos_t, sin_t, H, hd, kc3, vc3,
                                   page_size=page, paired=True)
    assert torch.equal(q3, q2) and torch.equal(kc3, kc2) and torch.equal(vc3, vc2)
    # fp32 restatement of HF apply_rotary_pos_emb on the bf16 projections
    qf = qkv[:, :D].view(M, H, hd)
    c, s = cos_t[positions.long()][:, None, :], sin_t[positions.long()][:, None, :]
    rot = torch.cat((-qf[..., hd // 2:], qf[..., : hd // 2]), -1)
    assert torch.equal(q.view(M, H, hd), (qf * c) + (rot * s))


def test_fused_decode_step_equals_unfused_chain(ctx):
    """The whole tiny model: greedy / scripted decoding with the fused 5-launch layers against the 9-launch chain."""
    from interactvlm_b200 import synthetic as S
    from interactvlm_b200.config import IVLMConfig
    from interactvlm_b200.model import InteractVLMForCausalLM
    from oracle.make_goldens_model import TINY_SEED, tiny_inputs

    cfg = IVLMConfig.tiny()
    sd = S.make_state_dict(cfg, seed=TINY_SEED["weights"])
    model = InteractVLMForCausalLM(cfg, sd, ctx=ctx)
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
    outs = {}
    for fused in (True, False):
        model.eng.fused_decode = fused
        model._graphs = {}
        n0 = ctx.launch_count()
        out_ids, hid = model.generate(clip, ids, max_new_tokens=ans.shape[1], scripted=ans)
        outs[fused] = (out_ids.clone(), hid.clone(), ctx.launch_count() - n0)
    model.eng.fused_decode = True
    assert torch.equal(outs[True][0], outs[False][0])
    n = outs[True][0].shape[1] - 1 + cfg.img_emb_len
    assert rel_err(outs[True][1][:, :n], outs[False][1][:, :n]) < 1e-2     # different GEMM kernels: bf16-level differences only
    assert outs[True][2] < outs[False][2]


@pytest.mark.parametrize("nxt_shape,stages", [((5120, 5120), 64), ((27648, 5120), 0), ((5120, 13824), 0), ((40, 256), 3), ((2000, 1088), 1000)])
def test_l2_prefetch_of_the_successor_does_not_change_results(ctx, nxt_shape, stages):
    """prefetch_w only issues cp.async.bulk.prefetch.L2 requests: outputs are bit-identical with it, without it and with the
    option that disables it, for successor shapes with ragged tiles / stages and fewer tiles than SMs."""
    M, N, K = 8, 5120, 5120
    a, w, res = rnd(M, K, seed=21), rnd(N, K, scale=K ** -0.5, seed=22), rnd(M, N, seed=23)
    nxt = rnd(*nxt_shape, scale=0.02, seed=24)
    base = ctx.decode_linear(a, w, residual=res)
    for kb in (-1, 64, 0):                                   # as asked / capped / off (the default)
        ctx.set_option("ds_prefetch_kb", kb)
        assert torch.equal(ctx.decode_linear(a, w, residual=res, prefetch=nxt, prefetch_stages=stages), base)
    torch.cuda.synchronize()


def _chain_case(ctx, M, D, F, H, hd, seed):
    """o_proj -> gate/up -> down_proj -> next qkv, once as four launches and once as one chained launch."""
    page, pages_per = 16, 4
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    fr = torch.outer(torch.arange(64).float(), inv)
    emb = torch.cat((fr, fr), -1)
    cos_t, sin_t = emb.cos().bfloat16().to(DEV), emb.sin().bfloat16().to(DEV)
    o, x = rnd(M, D, seed=seed), rnd(M, D, seed=seed + 1)
    wo, wd = rnd(D, D, scale=D ** -0.5, seed=seed + 2), rnd(D, F, scale=F ** -0.5, seed=seed + 3)
    wgu = interleave_gate_up(rnd(F, D, scale=D ** -0.5, seed=seed + 4), rnd(F, D, scale=D ** -0.5, seed=seed + 5))
    wq, wk, wv = (rnd(D, D, scale=D ** -0.5, seed=seed + s) for s in (6, 7, 8))
    wqkv = torch.cat([pair_rows(wq, H, hd), pair_rows(wk, H, hd), wv], 0)
    g1 = (1 + 0.1 * torch.randn(D, generator=torch.Generator().manual_seed(seed + 9))).bfloat16().to(DEV)
    g2 = (1 + 0.1 * torch.randn(D, generator=torch.Generator().manual_seed(seed + 10))).bfloat16().to(DEV)
    positions = torch.arange(M, dtype=torch.int32, device=DEV) * 5 + 3
    slot = (torch.arange(M, dtype=torch.int32, device=DEV) * pages_per * page + positions).contiguous()
    res = {}
    for chained in (False, True):
        kc = torch.zeros(M * pages_per, H, page, hd, device=DEV, dtype=torch.bfloat16)
        vc = torch.zeros_like(kc)
        rope = dict(positions=positions, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc, v_cache=vc, H=H, hd=hd, page_size=page)
        mk = lambda w: torch.full((M, w), 7.0, device=DEV, dtype=torch.bfloat16)
        xn, act, x2, q = mk(D), mk(F), mk(D), mk(D)
        phases = [(o, wo, dict(residual=x, out=xn)),
                  (xn, wgu, dict(gamma=g2, eps=1e-5, epilogue=1, out=act)),
                  (act, wd, dict(residual=xn, out=x2)),
                  (x2, wqkv, dict(gamma=g1, eps=1e-5, epilogue=2, rope=rope, out=q))]
        if chained:
            for _ in range(3):                       # barriers re-arm themselves between launches
                ctx.decode_chain(phases)
        else:
            for a_, w_, kw in phases:
                ctx.decode_linear(a_, w_, **kw)
        torch.cuda.synchronize()
        res[chained] = [t.clone() for t in (xn, act, x2, q, kc, vc)]
    for name, a_, b_ in zip(("o_proj", "gate_up", "down", "q", "k_cache", "v_cache"), res[False], res[True]):
        assert torch.equal(a_, b_), name
    ref = (o.float() @ wo.float().t()).bfloat16().float() + x.float()
    assert rel_err(res[True][0], ref) < 4e-3


@pytest.mark.parametrize("M,D,F,H,hd", [(8, 5120, 13824, 40, 128), (3, 256, 704, 4, 64), (8, 512, 1408, 4, 128), (1, 128, 320, 2, 64)])
def test_chained_launch_equals_separate_launches(ctx, M, D, F, H, hd):
    """ivlm_decode_chain against four ivlm_decode_linear launches: bit-identical outputs (13B widths; tiny widths where phases have
    fewer tiles than SMs, ragged K stages and M < 8), repeated launches (self-re-arming grid barriers)."""
    _chain_case(ctx, M, D, F, H, hd, seed=100)


def test_chained_launch_shorter_chains(ctx):
    M, D, F = 4, 1024, 2816
    o, x = rnd(M, D, seed=1), rnd(M, D, seed=2)
    wo, wd = rnd(D, D, scale=D ** -0.5, seed=3), rnd(D, F, scale=F ** -0.5, seed=4)
    y = rnd(M, F, seed=5)
    a1 = ctx.decode_linear(o, wo, residual=x)
    b1 = ctx.decode_linear(y, wd, residual=a1)
    (a2,) = ctx.decode_chain([(o, wo, dict(residual=x))])
    assert torch.equal(a1, a2)
    a3 = torch.empty_like(a1)
    outs = ctx.decode_chain([(o, wo, dict(residual=x, out=a3)), (y, wd, dict(residual=a3))])
    assert torch.equal(outs[0], a1) and torch.equal(outs[1], b1)


def test_chained_decode_step_is_bit_identical(ctx):
    """The tiny model with model.eng.chained_decode (attention + one ivlm_decode_chain launch per layer) against the default
    five launches per layer: same tokens, identical hidden states."""
    from interactvlm_b200 import synthetic as S
    from interactvlm_b200.config import IVLMConfig
    from interactvlm_b200.model import InteractVLMForCausalLM
    from oracle.make_goldens_model import TINY_SEED, tiny_inputs

    cfg = IVLMConfig.tiny()
    model = InteractVLMForCausalLM(cfg, S.make_state_dict(cfg, seed=TINY_SEED["weights"]), ctx=ctx)
    ids, ans, clip, sam, cam = tiny_inputs(cfg, 2)
    outs = {}
    model.eng.stage_abi = False          # the op-level decode step (the C stage driver always issues the five launches)
    for chained in (False, True):
        model.eng.chained_decode = chained
        model._graphs = {}
        n0 = ctx.launch_count()
        out_ids, hid = model.generate(clip, ids, max_new_tokens=ans.shape[1], scripted=ans)
        outs[chained] = (out_ids.clone(), hid.clone(), ctx.launch_count() - n0)
    model.eng.chained_decode = False
    model.eng.stage_abi = True
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    assert outs[True][2] < outs[False][2]
