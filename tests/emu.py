"""TEST-ONLY stand-in for interactvlm_b200.ops.Context on a machine without a GPU.

Each method states, in plain torch on CPU, what the corresponding sm_100a kernel computes (same argument meaning,
same layouts, bf16 storage with fp32 arithmetic).  It exists so the HOST logic of interactvlm_b200/model.py -- stage
order, weight re-layout, row maps, KV-cache bookkeeping, [SEG] indexing -- can be checked against the oracle in the
CPU test tier.  It is never imported by the product package and is not a fallback: ops.Context raises without CUDA.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BF = torch.bfloat16
ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, ACT_SWIGLU = 0, 1, 2, 3, 4, 5


def _act(x, act):
    if act == ACT_GELU:
        return F.gelu(x)
    if act == ACT_QUICK_GELU:
        return x * torch.sigmoid(1.702 * x)
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_SILU:
        return F.silu(x)
    return x


def _r(x):
    return x.to(BF).float()


class EmuLiftMap:
    def __init__(self, p2v, bary, n):
        self.p2v, self.bary, self.n = np.asarray(p2v), (None if bary is None else np.asarray(bary)), n
        self.V, self.H, self.W = self.p2v.shape[:3]

    def __call__(self, masks, mode, thr=0.3):
        from oracle import lift as OL

        m = masks.numpy()
        if mode == 0:
            return torch.from_numpy(OL.lift_human(m, self.p2v, self.bary, self.n))
        if mode == 1:
            return torch.from_numpy(OL.lift_object_mesh(m, self.p2v, self.bary, self.n, thr))
        return torch.from_numpy(OL.lift_points(m, self.p2v, self.n))


    def lowres(self, low, mode, thr=0.3):
        """ivlm_lift_lowres: bilinear to the map size, then the lift (what the fused kernel computes)."""
        import torch.nn.functional as F

        H, W = self.p2v.shape[1], self.p2v.shape[2]
        return self(F.interpolate(low.float(), (H, W), mode="bilinear", align_corners=False), mode, thr)


class EmuContext:
    emulated = True

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def launch_count(self):
        return self.launches

    def LiftMap(self, p2v, bary, n):
        return EmuLiftMap(p2v, bary, n)

    # ---- dense
    def gemm(self, a, w, bias=None, act=ACT_NONE, residual=None, out=None, out_dtype=BF, row_map=None, out_rows=None,
             k_splits=1, force_swap=0, no_round=False, res_row_mod=0):
        self.launches += 1
        assert a.dtype == BF and w.dtype == BF and a.shape[1] == w.shape[1] and a.shape[1] % 8 == 0
        if act == ACT_SWIGLU:   # interleaved gate / up rows -> bf16(bf16(silu(gate)) * up), N / 2 columns (the staged GEMM epilogue)
            assert a.shape[0] > 64 and bias is None and residual is None and row_map is None
            y = self.silu_mul(_r(a.float() @ w.float().t()).to(BF), interleaved=True)
            self.launches -= 1
            if out is not None:
                out.copy_(y)
                return out
            return y
        M, N = a.shape[0], w.shape[0]
        swap = force_swap == 1 or (force_swap == 0 and M <= 64 and row_map is None)
        if not swap:
            assert N % 8 == 0, "row-major epilogue needs N % 8 == 0"
        to_bf = (out.dtype if out is not None else out_dtype) == BF
        rnd = _r if (to_bf and not no_round) else (lambda t: t)
        y = a.float() @ w.float().t()
        if bias is not None:
            y = y + bias.float()
        y = rnd(y)
        if act != ACT_NONE:
            y = rnd(_act(y, act))
        if out is None:
            rows = M if out_rows is None else out_rows
            out = torch.zeros((rows, N), dtype=out_dtype)
        if row_map is not None:
            rm = row_map.long()
            live = rm >= 0
            orow = rm[live]
            y = y[live]
        else:
            orow = torch.arange(M)
        if residual is not None:
            rrow = orow % res_row_mod if res_row_mod > 0 else orow
            y = y + residual.float()[rrow]
        out[orow] = y.to(out.dtype)
        return out

    # ---- norms / elementwise
    def layernorm(self, x, gamma, beta, eps, row_map=None, out_rows=None, act=ACT_NONE, out=None):
        self.launches += 1
        D = x.shape[-1]
        x2 = x.reshape(-1, D).float()
        if row_map is not None:
            rm = row_map.long()
            src = x2[rm.clamp(min=0)]
        else:
            src = x2
        y = F.layer_norm(src, (D,), gamma.float(), beta.float(), eps)
        if act != ACT_NONE:
            y = _act(_r(y), act)
        y = y.to(BF)
        if row_map is not None:
            y[rm < 0] = 0
            return y
        return y.view(x.shape)

    def rmsnorm(self, x, gamma, eps, out=None):
        self.launches += 1
        xf = x.float()
        y = (gamma.float() * _r(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(BF)
        if out is not None:
            out.copy_(y.view(out.shape))
            return out.view(x.shape)
        return y

    def add_bcast(self, a, b, out=None):
        self.launches += 1
        assert a.is_contiguous() and b.is_contiguous() and a.numel() % b.numel() == 0
        return (a.float().view(-1, b.numel()) + b.float().view(1, -1)).to(BF).view(a.shape)

    def silu_mul(self, gate_up, out=None, interleaved=False):
        self.launches += 1
        Fh = gate_up.shape[1] // 2
        if interleaved:
            from interactvlm_b200.layout import split_interleaved
            gate, up = split_interleaved(gate_up)
        else:
            gate, up = gate_up[:, :Fh], gate_up[:, Fh:]
        return (_r(F.silu(gate.float())) * up.float()).to(BF)

    def decode_chain(self, phases):
        """ivlm_decode_chain: the phases one after the other; each writes its `out` tensor, which later phases read."""
        return [self.decode_linear(a, w, **kw) for a, w, kw in phases]

    def decode_linear(self, a, w, gamma=None, eps=0.0, epilogue=0, act=ACT_NONE, bias=None, residual=None, out=None, out_dtype=BF,
                      rope=None, prefetch=None, prefetch_stages=0):
        """ivlm_decode_linear: rmsnorm? -> linear -> PLAIN / SWIGLU (interleaved rows) / ROPE_KV (paired q, k rows)."""
        assert a.shape[0] <= 8 and a.shape[1] % 16 == 0
        x = self.rmsnorm(a, gamma, eps) if gamma is not None else a
        if epilogue == 0:
            y = self.gemm(x, w, bias=bias, act=act, residual=residual, out_dtype=out_dtype, force_swap=1)
        elif epilogue == 1:
            y = self.silu_mul(self.gemm(x, w, force_swap=1), interleaved=True)
        else:
            qkv = self.gemm(x, w, force_swap=1)
            y, _, _ = self.rope_kv_store(qkv, rope["positions"], rope["slot_map"], rope["cos"], rope["sin"], rope["H"], rope["hd"],
                                         rope["k_cache"], rope["v_cache"], want_kv=False, page_size=rope["page_size"], paired=True)
        if out is not None:
            out.copy_(y)
            return out
        return y

    # ---- lowering
    def im2col_patch(self, img, p, ldk=None):
        self.launches += 1
        N, C, H, W = img.shape
        kk = C * p * p
        ldk = ldk or (kk + 7) // 8 * 8
        cols = F.unfold(img.float(), p, stride=p).transpose(1, 2).reshape(-1, kk)
        out = torch.zeros((cols.shape[0], ldk), dtype=BF)
        out[:, :kk] = cols.to(BF)
        return out

    def im2col_3x3(self, x, N, H, W):
        self.launches += 1
        C = x.shape[-1]
        t = x.float().view(N, H, W, C).permute(0, 3, 1, 2)
        u = F.unfold(t, 3, padding=1).view(N, C, 9, H * W).permute(0, 3, 2, 1).reshape(N * H * W, 9 * C)
        return u.to(BF)

    # ---- attention
    def attention(self, q, k, v, scale, causal=False, rel_h=None, rel_w=None, kh=0, kw=0, out=None):
        self.launches += 1
        B, Sq, H, D = q.shape
        Sk = k.shape[1]
        qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
        s = qf @ kf.transpose(-1, -2) * scale
        if rel_h is not None:
            s = (s.view(B, H, Sq, kh, kw) + rel_h.view(B, H, Sq, kh, 1) + rel_w.view(B, H, Sq, 1, kw)).view(B, H, Sq, Sk)
        if causal:
            s = s.masked_fill(~torch.ones(Sq, Sk).tril(Sk - Sq).bool(), float("-inf"))
        return (s.softmax(-1) @ vf).permute(0, 2, 1, 3).to(BF).contiguous()

    def sam_relpos(self, qkv, rel_pos_h, rel_pos_w, B, heads, Hq, Wq, hd):
        self.launches += 1
        S = Hq * Wq
        q = qkv.view(B, S, 3, heads, hd)[:, :, 0].float().permute(0, 2, 1, 3).reshape(B, heads, Hq, Wq, hd)
        ih = torch.arange(Hq)[:, None] - torch.arange(Hq)[None, :] + Hq - 1
        iw = torch.arange(Wq)[:, None] - torch.arange(Wq)[None, :] + Wq - 1
        Rh, Rw = rel_pos_h.float()[ih], rel_pos_w.float()[iw]
        rel_h = _r(torch.einsum("bnhwc,hkc->bnhwk", q, Rh)).reshape(B, heads, S, Hq)
        rel_w = _r(torch.einsum("bnhwc,wkc->bnhwk", q, Rw)).reshape(B, heads, S, Wq)
        return rel_h.contiguous(), rel_w.contiguous()

    def sam_attention(self, qkv, rel_pos_h, rel_pos_w, B, heads, Hq, Wq, hd, out=None, out_map=None, out_rows=None):
        rel_h, rel_w = self.sam_relpos(qkv, rel_pos_h, rel_pos_w, B, heads, Hq, Wq, hd)
        t = qkv.view(B, Hq * Wq, 3, heads, hd)
        o = self.attention(t[:, :, 0], t[:, :, 1], t[:, :, 2], hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=Hq, kw=Wq)
        o = o.reshape(B * Hq * Wq, heads * hd)
        if out_map is None:
            return o
        res = torch.zeros((out_rows, heads * hd), dtype=BF)
        live = out_map.long() >= 0
        res[out_map.long()[live]] = o[live]
        return res

    def fill_rows(self, out, rows, vec):
        self.launches += 1
        out[rows.long()] = vec
        return out

    def attn_small(self, q, k, v, heads):
        self.launches += 1
        B, Nk, C = k.shape
        Nq = q.shape[1]
        hd = C // heads
        assert Nk <= 16 or Nq <= 64
        qf = q.float().expand(B, Nq, C).reshape(B, Nq, heads, hd).transpose(1, 2)
        kf = k.float().reshape(B, Nk, heads, hd).transpose(1, 2)
        vf = v.float().reshape(B, Nk, heads, hd).transpose(1, 2)
        a = _r(_r(qf @ kf.transpose(-1, -2)) / math.sqrt(hd)).softmax(-1)
        return (_r(a) @ vf).transpose(1, 2).reshape(B, Nq, C).to(BF)

    # ---- LLaVA / LLaMA glue
    def embed_splice(self, embed, ids, img_feats):
        self.launches += 1
        rows = []
        for b in range(ids.shape[0]):
            r = ids[b].long()
            pos = int((r < 0).nonzero()[0])
            rows.append(torch.cat([embed[r[:pos]], img_feats[b], embed[r[pos + 1:]]], 0))
        return torch.stack(rows, 0)

    def embed_gather(self, embed, ids, out=None):
        self.launches += 1
        return embed[ids.long()]

    def gather_rows(self, x, idx):
        self.launches += 1
        return x[idx.long()]

    def rows_differ(self, x, ref):
        self.launches += 1
        a, b = x.reshape(x.shape[0], -1), ref.reshape(ref.shape[0], -1)
        return (a.view(torch.int16)[:, None, :] != b.view(torch.int16)[None, :, :]).any(-1).to(torch.int32)

    def rope_kv_store(self, qkv, positions, slot_map, cos_t, sin_t, H, hd, k_cache=None, v_cache=None, want_kv=True,
                      q_out=None, page_size=16, paired=False):
        self.launches += 1
        T = qkv.shape[0]
        D = H * hd
        if paired:   # q and k columns arrive in the paired order of layout.py
            from interactvlm_b200.layout import unpair_cols
            qkv = torch.cat([unpair_cols(qkv[:, :D], H, hd), unpair_cols(qkv[:, D:2 * D], H, hd), qkv[:, 2 * D:]], 1)
        q, k, v = (qkv[:, i * D:(i + 1) * D].reshape(T, H, hd) for i in range(3))
        c, s = cos_t[positions.long()][:, None, :], sin_t[positions.long()][:, None, :]

        def rope(x):  # bf16 ops like HF apply_rotary_pos_emb
            h = hd // 2
            rot = torch.cat((-x[..., h:], x[..., :h]), -1)
            return (x * c) + (rot * s)

        qo, ko = rope(q), rope(k)
        if slot_map is not None:  # cache layout [pages, H, page, hd]
            sl = slot_map.long()
            k_cache[sl // page_size, :, sl % page_size] = ko
            v_cache[sl // page_size, :, sl % page_size] = v
        return qo.reshape(T, D), (ko.reshape(T, D) if want_kv else None), (v.reshape(T, D).contiguous() if want_kv else None)

    def decode_attention(self, q, k_cache, v_cache, block_table, seq_lens, H, hd, page_size, out=None):
        self.launches += 1
        B = q.shape[0]
        res = torch.empty((B, H * hd), dtype=BF)
        for b in range(B):
            n = int(seq_lens[b])
            pos = torch.arange(n)
            pg, off = block_table[b].long()[pos // page_size], pos % page_size
            kf, vf = k_cache[pg, :, off].float(), v_cache[pg, :, off].float()  # [n,H,hd]
            sc = _r(_r(torch.einsum("hd,nhd->hn", q[b].view(H, hd).float(), kf)) / math.sqrt(hd))
            p = _r(sc.softmax(-1))
            res[b] = torch.einsum("hn,nhd->hd", p, vf).reshape(-1).to(BF)
        return res

    def decode_prepare(self, st, S, G, eos, pad):
        self.launches += 1
        step = int(st["state"][0])
        scripted = st.get("scripted")
        t = (scripted[:, step] if scripted is not None else st["next"]).clone().to(torch.int32)
        t[st["done"].bool()] = pad
        st["out_tokens"][:, step] = t
        st["done"][t == eos] = 1
        rows = st.get("S_rows")
        p = (rows.to(torch.int32) if rows is not None else torch.full_like(t, S)) + step
        st["tok"].copy_(t)
        st["pos"].copy_(p)
        st["slot"].copy_(st["slot_base"] + p)
        st["seq_lens"].copy_(p + 1)
        st["state"][1] = step
        st["state"][0] = step + 1

    def decode_finish(self, st, S):
        self.launches += 1
        rows = st.get("S_rows")
        step = int(st["state"][1])
        for b in range(st["hidden"].shape[0]):
            st["hidden"][b, (int(rows[b]) if rows is not None else S) + step] = st["hid_step"][b]

    def argmax(self, logits, vocab=None, out=None):
        self.launches += 1
        r = logits[:, :vocab].argmax(-1).to(torch.int32)
        if out is not None:
            out.copy_(r)
            return out
        return r

    # ---- prompt / mask tail
    def cam_gate(self, cam, emb, w1, b1, w2, b2, wv, bv):
        self.launches += 1
        h = _r(F.relu(cam.float() @ w1.float().t() + b1.float()))
        h = _r(F.relu(h @ w2.float().t() + b2.float()))
        g = _r(torch.sigmoid(_r(torch.einsum("bvk,vnk->bvn", h, wv.float()) + bv.float())))
        return (emb.float()[:, None, :] * g).to(BF)

    def upscale_hyper_dot(self, up1, w2, b2, hyper, Bv, grid):
        self.launches += 1
        G = grid
        u = up1.float().view(Bv, G * G, 4, -1)
        z = _r(torch.einsum("btpk,qck->btpqc", u, w2.float()) + b2.float())
        z = _r(F.gelu(z))
        m = torch.einsum("btpqc,bc->btpq", z, hyper.float())
        m = m.view(Bv, G, G, 2, 2, 2, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(Bv, G * 4, G * 4)
        return _r(m)

    def preprocess_u8(self, img_u8, size, kind="sam"):
        self.launches += 1
        from interactvlm_b200.ops import Context as _C
        mean, std, pre = (_C.SAM_MEAN, _C.SAM_STD, 1.0) if kind == "sam" else (_C.CLIP_MEAN, _C.CLIP_STD, 1.0 / 255.0)
        x = img_u8.float().permute(0, 3, 1, 2) * pre
        x = (x - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
        N, _, H, W = x.shape
        return F.pad(x, (0, size - W, 0, size - H)).to(BF)

    def resize_u8(self, img_u8, out_h, out_w, filt="bilinear"):
        self.launches += 1
        from oracle import resize as OR
        return torch.stack([torch.from_numpy(OR.resize_u8(i.numpy(), out_h, out_w, filt)) for i in img_u8], 0)

    def sigmoid_where(self, x, gt=None, ignore_value=-1.0):
        self.launches += 1
        keep = torch.ones_like(x, dtype=torch.bool) if gt is None else gt != ignore_value
        x[keep] = torch.sigmoid(x[keep])
        return x

    def bilinear(self, src, dh, dw, crop_h=None, crop_w=None, out=None):
        self.launches += 1
        ch, cw = crop_h or src.shape[1], crop_w or src.shape[2]
        return F.interpolate(src[None, :, :ch, :cw], (dh, dw), mode="bilinear", align_corners=False)[0].contiguous()
