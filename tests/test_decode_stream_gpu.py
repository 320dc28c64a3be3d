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
