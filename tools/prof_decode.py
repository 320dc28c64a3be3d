"""GPU-box helper: per-op time of one LLaMA-13B decode layer at batch 8 inside CUDA graphs (no launch gaps from Python),
weights rotated over 4 copies so that L2 cannot hold them.  Prints us per op and the achieved HBM GB/s."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from interactvlm_b200.ops import Context  # noqa: E402

ctx = Context(0)
import os  # noqa: E402

B, D, F, H, hd, L = int(os.environ.get("PROF_B", "8")), 5120, 13824, 40, 128, 340
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).bfloat16()
NW = 4
wqkv = [rnd(3 * D, D, sc=0.02) for _ in range(NW)]
wo = [rnd(D, D, sc=0.02) for _ in range(NW)]
wgu = [rnd(2 * F, D, sc=0.02) for _ in range(NW)]
wd = [rnd(D, F, sc=0.02) for _ in range(NW)]
x, xf, gam = rnd(B, D), rnd(B, F), rnd(D)
gu = rnd(B, 2 * F)
qkv = rnd(B, 3 * D)
page, pages = 16, 24
kc = [rnd(B * pages, H, page, hd) for _ in range(NW)]
vc = [rnd(B * pages, H, page, hd) for _ in range(NW)]
bt = torch.arange(B * pages, dtype=torch.int32, device="cuda").view(B, pages)
sl = torch.full((B,), L, dtype=torch.int32, device="cuda")
pos = torch.full((B,), L - 1, dtype=torch.int32, device="cuda")
slot = (torch.arange(B, dtype=torch.int32, device="cuda") * pages * page + L - 1).contiguous()
inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
fr = torch.outer(torch.arange(1024).float(), inv)
emb = torch.cat((fr, fr), -1)
cos_t, sin_t = emb.cos().bfloat16().cuda(), emb.sin().bfloat16().cuda()
q = rnd(B, D)
REP = 40


def graph_time(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(REP):
            fn(i)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / REP * 1e3  # us per op


ops = {
    "gemm_qkv": (lambda i: ctx.gemm(x, wqkv[i % NW]), 3 * D * D * 2),
    "gemm_o": (lambda i: ctx.gemm(x, wo[i % NW], residual=x), D * D * 2),
    "gemm_gateup": (lambda i: ctx.gemm(x, wgu[i % NW]), 2 * F * D * 2),
    "gemm_down": (lambda i: ctx.gemm(xf, wd[i % NW], residual=x), F * D * 2),
    "gemm_o_nosplit": (lambda i: ctx.gemm(x, wo[i % NW], residual=x, k_splits=1), D * D * 2),
    "gemm_down_nosplit": (lambda i: ctx.gemm(xf, wd[i % NW], residual=x, k_splits=1), F * D * 2),
    "rmsnorm": (lambda i: ctx.rmsnorm(x, gam, 1e-5), 0),
    "rope_kv_store": (lambda i: ctx.rope_kv_store(qkv, pos, slot, cos_t, sin_t, H, hd, kc[i % NW], vc[i % NW], want_kv=False, page_size=page), 0),
    "decode_attention": (lambda i: ctx.decode_attention(q, kc[i % NW], vc[i % NW], bt, sl, H, hd, page), 2 * B * L * D * 2),
    "silu_mul": (lambda i: ctx.silu_mul(gu), 0),
}
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    gemms = {k: v for k, v in ops.items() if k.startswith("gemm") and "nosplit" not in k}
    cfgs = [("tcgen05 swapped + split-K", dict(small_m_variant=1))]
    for rows8 in (0, 1):
        for wv in (3, 5, 6, 7, 10, 12, 16):
            cfgs.append((f"gemv rows{8 if rows8 else 16} warps{wv}", dict(small_m_variant=0, gv_rows8_max_n=(1 << 30) if rows8 else 0, gv_warps=wv)))
    for label, opts in cfgs:
        for k, v in opts.items():
            ctx.set_option(k, v)
        print(label + ": " + "  ".join(f"{n[5:]} {graph_time(fn):6.2f}us" for n, (fn, _) in gemms.items()), flush=True)
    sys.exit(0)
def layer_chain(i):
    """the 9 launches of one decode layer, in order, on rotating weights (what the captured decode graph replays)"""
    y = ctx.rmsnorm(x, gam, 1e-5)
    qkv_ = ctx.gemm(y, wqkv[i % NW])
    q_, _, _ = ctx.rope_kv_store(qkv_, pos, slot, cos_t, sin_t, H, hd, kc[i % NW], vc[i % NW], want_kv=False, page_size=page)
    o_ = ctx.decode_attention(q_, kc[i % NW], vc[i % NW], bt, sl, H, hd, page)
    x1 = ctx.gemm(o_, wo[i % NW], residual=x)
    y = ctx.rmsnorm(x1, gam, 1e-5)
    y = ctx.silu_mul(ctx.gemm(y, wgu[i % NW]))
    return ctx.gemm(y, wd[i % NW], residual=x1)


def fused_chain(i):
    """the 5 launches of one decode layer with ivlm_decode_linear (RMSNorm / RoPE + KV store / SwiGLU fused)"""
    rope = dict(positions=pos, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc[i % NW], v_cache=vc[i % NW], H=H, hd=hd, page_size=page)
    q_ = ctx.decode_linear(x, wqkv[i % NW], gamma=gam, eps=1e-5, epilogue=2, rope=rope)
    o_ = ctx.decode_attention(q_, kc[i % NW], vc[i % NW], bt, sl, H, hd, page)
    x1 = ctx.decode_linear(o_, wo[i % NW], residual=x)
    y = ctx.decode_linear(x1, wgu[i % NW], gamma=gam, eps=1e-5, epilogue=1)
    return ctx.decode_linear(y, wd[i % NW], residual=x1)


def fused_chain_pf(i, wo_stages=64, stages=0):
    """fused_chain with every launch requesting the head of its successor's weights into L2 (as model.py drives it)"""
    rope = dict(positions=pos, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc[i % NW], v_cache=vc[i % NW], H=H, hd=hd, page_size=page)
    q_ = ctx.decode_linear(x, wqkv[i % NW], gamma=gam, eps=1e-5, epilogue=2, rope=rope, prefetch=wo[i % NW], prefetch_stages=wo_stages)
    o_ = ctx.decode_attention(q_, kc[i % NW], vc[i % NW], bt, sl, H, hd, page)
    x1 = ctx.decode_linear(o_, wo[i % NW], residual=x, prefetch=wgu[i % NW], prefetch_stages=stages)
    y = ctx.decode_linear(x1, wgu[i % NW], gamma=gam, eps=1e-5, epilogue=1, prefetch=wd[i % NW], prefetch_stages=stages)
    return ctx.decode_linear(y, wd[i % NW], residual=x1, prefetch=wqkv[(i + 1) % NW], prefetch_stages=stages)


xn_b, act_b, x_b, q_b = torch.empty_like(x), torch.empty_like(xf), torch.empty_like(x), torch.empty_like(x)


def chained_chain(i):
    """attention + ONE chained launch (o_proj -> gate/up -> down_proj -> next qkv), as model.py drives a layer"""
    rope = dict(positions=pos, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc[(i + 1) % NW], v_cache=vc[(i + 1) % NW], H=H, hd=hd, page_size=page)
    o_ = ctx.decode_attention(q_b, kc[i % NW], vc[i % NW], bt, sl, H, hd, page)
    ctx.decode_chain([(o_, wo[i % NW], dict(residual=x, out=xn_b)),
                      (xn_b, wgu[i % NW], dict(gamma=gam, eps=1e-5, epilogue=1, out=act_b)),
                      (act_b, wd[i % NW], dict(residual=xn_b, out=x_b)),
                      (x_b, wqkv[(i + 1) % NW], dict(gamma=gam, eps=1e-5, epilogue=2, rope=rope, out=q_b))])
    return q_b


if len(sys.argv) > 1 and sys.argv[1] == "chain2":
    for pdl in (0, 1):
        ctx.set_option("pdl", pdl)
        print(f"pdl={pdl}: 5-launch layer {graph_time(fused_chain):7.2f} us;   attention + chained launch {graph_time(chained_chain):7.2f} us", flush=True)
    ctx.set_option("pdl", 1)
    for ns in (5, 4, 3):
        ctx.set_option("ds_stages", ns)
        print(f"chained launch, ring depth {ns}: {graph_time(chained_chain):7.2f} us per layer", flush=True)
    ctx.set_option("ds_stages", 0)
    ctx.set_option("pdl", 0)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "prefetch":
    ctx.set_option("pdl", 1)
    ctx.set_option("ds_prefetch_kb", -1)   # the prefetch is off by default
    ref = fused_chain(0).float()
    print(f"fused chain, no prefetch: {graph_time(fused_chain):7.2f} us per layer", flush=True)
    for wo_st, st in ((64, 0), (64, 8), (64, 32), (64, 64), (16, 16), (0, 16), (64, 4)):
        us = graph_time(lambda i: fused_chain_pf(i, wo_st, st))
        same = bool(torch.equal(fused_chain_pf(0, wo_st, st).float(), ref))
        print(f"fused chain, L2 prefetch of the successor: wo {wo_st or 16} stages/SM, others {st or 16} stages/SM: {us:7.2f} us per layer; "
              f"identical output: {same}", flush=True)
    ctx.set_option("ds_prefetch_kb", 0)
    ctx.set_option("pdl", 0)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "fused":
    fops = {
        "norm+qkv+rope+kv": (lambda i: ctx.decode_linear(x, wqkv[i % NW], gamma=gam, eps=1e-5, epilogue=2, rope=dict(
            positions=pos, slot_map=slot, cos=cos_t, sin=sin_t, k_cache=kc[i % NW], v_cache=vc[i % NW], H=H, hd=hd, page_size=page)), 3 * D * D * 2),
        "o_proj+res": (lambda i: ctx.decode_linear(x, wo[i % NW], residual=x), D * D * 2),
        "norm+gateup+swiglu": (lambda i: ctx.decode_linear(x, wgu[i % NW], gamma=gam, eps=1e-5, epilogue=1), 2 * F * D * 2),
        "down+res": (lambda i: ctx.decode_linear(xf, wd[i % NW], residual=x), F * D * 2),
        "decode_attention": (lambda i: ctx.decode_attention(q, kc[i % NW], vc[i % NW], bt, sl, H, hd, page), 2 * B * L * D * 2),
    }
    for name, (fn, nbytes) in fops.items():
        us = graph_time(fn)
        print(f"{name:>20}: {us:8.2f} us   {nbytes / us / 1e3:8.1f} GB/s", flush=True)
    for pdl in (0, 1):
        ctx.set_option("pdl", pdl)
        print(f"fused layer chain in a graph, pdl={pdl}: {graph_time(fused_chain):7.2f} us per layer;   9-launch chain: {graph_time(layer_chain):7.2f} us", flush=True)
    ctx.set_option("pdl", 1)
    for ns in (7, 6, 5, 4, 3):   # 8 stages no longer fit next to the resident activation + staged RMSNorm weight
        ctx.set_option("ds_stages", ns)
        row = "  ".join(f"{name.split('+')[0]} {graph_time(fn):6.2f}" for name, (fn, _) in list(fops.items())[:4])
        print(f"ring depth {ns}: {row}  chain {graph_time(fused_chain):7.2f} us", flush=True)
    ctx.set_option("ds_stages", 0)
    for wv in (8, 11, 16):
        ctx.set_option("dec_warps", wv)
        print(f"decode attention with {wv} warps per CTA: {graph_time(fops['decode_attention'][0]):6.2f} us   chain {graph_time(fused_chain):7.2f} us", flush=True)
    ctx.set_option("dec_warps", 0)
    for pf in (0, 1):
        ctx.set_option("dec_prefetch", pf)
        print(f"decode attention, K / V lines requested into L2 up front = {pf}: {graph_time(fops['decode_attention'][0]):6.2f} us   "
              f"chain {graph_time(fused_chain):7.2f} us", flush=True)
    for fs in (1, 0):
        ctx.set_option("ds_force_stream", fs)
        print(f"o_proj with the activation {'streamed' if fs else 'resident'}: {graph_time(fops['o_proj+res'][0]):6.2f} us   "
              f"chain {graph_time(fused_chain):7.2f} us", flush=True)
    ctx.set_option("pdl", 0)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "splits":
    ctx.set_option("small_m_variant", 1)
    for name, wl, xin in (("qkv", wqkv, x), ("o", wo, x), ("gateup", wgu, x), ("down", wd, xf)):
        row = []
        for ks in (0, 1, 2, 3, 4, 6):
            ctx.set_option("fused_split_force", ks)   # 0: the cost model's choice
            row.append(f"split {ks or 'auto'}: {graph_time(lambda i, wl=wl, xin=xin: ctx.gemm(xin, wl[i % NW])):6.2f}us")
        ctx.set_option("fused_split_force", 0)
        print(f"B={B} {name}: " + "  ".join(row), flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "variants":
    gemms = {k: v for k, v in ops.items() if k.startswith("gemm") and "nosplit" not in k}
    for label, v in (("gemv (N <= 8192) + swapped tcgen05", 0), ("swapped tcgen05 + split-K everywhere", 1)):
        ctx.set_option("small_m_variant", v)
        print(f"B={B} {label}: " + "  ".join(f"{n[5:]} {graph_time(fn):6.2f}us" for n, (fn, _) in gemms.items()), flush=True)
    ctx.set_option("small_m_variant", 0)
    print(f"B={B} decode_attention {graph_time(ops['decode_attention'][0]):6.2f}us rmsnorm {graph_time(ops['rmsnorm'][0]):5.2f}us "
          f"rope {graph_time(ops['rope_kv_store'][0]):5.2f}us silu {graph_time(ops['silu_mul'][0]):5.2f}us", flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "chain":
    ref = None
    for pdl in (0, 1, 0, 1):
        ctx.set_option("pdl", pdl)
        us = graph_time(layer_chain)
        out = layer_chain(0).float()
        torch.cuda.synchronize()
        same = True if ref is None else bool(torch.equal(out, ref))
        ref = out if ref is None else ref
        print(f"layer chain in a graph, pdl={pdl}: {us:7.2f} us per layer -> {us * 40 / 1e3:.2f} ms per 40-layer step; identical output: {same}", flush=True)
    ctx.set_option("pdl", 0)
    sys.exit(0)
tot = 0.0
for name, (fn, nbytes) in ops.items():
    us = graph_time(fn)
    print(f"{name:>20}: {us:8.2f} us" + (f"   {nbytes / us / 1e3:8.1f} GB/s" if nbytes else ""), flush=True)
    if "nosplit" not in name:
        tot += us * (2 if name == "rmsnorm" else 1)
print(f"layer total {tot:.1f} us -> x40 = {tot * 40 / 1e3:.2f} ms per decode step (ideal weights-only at 6551 GB/s: {12 * D * D * 2 / 6551e3 * 40 / 1e3 + 0:.2f} ms)")
