"""Torch-tensor front end of the C ABI.  PyTorch is plumbing only: it owns device memory and streams;
every computation below is one of our sm_100a kernels launched through libivlm_b200.so."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, BF16, F32, LIFT_HUMAN,  # noqa: F401
                   LIFT_OBJECT_MESH, LIFT_POINTS)

i32, i64, f32c = C.c_int32, C.c_int64, C.c_float


def P(t):
    if t is None:
        return C.c_void_p(None)
    return C.c_void_p(t.data_ptr())


def _bf16(t, name="tensor"):
    assert t.dtype == torch.bfloat16 and t.is_cuda, f"{name}: expected a CUDA bf16 tensor, got {t.dtype} on {t.device}"
    return t


class Context:
    """One ivlm handle bound to one CUDA device."""

    def __init__(self, device: int | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("interactvlm_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self.lib = L.lib()
        h = C.c_void_p()
        torch.cuda.set_device(self.device)
        L.check(self.lib.ivlm_create(C.byref(h), i32(self.device.index)), "ivlm_create")
        self.h = h
        self.profiling = False   # per-op event brackets on (enable_profile): stage-level calls are then bypassed by the model
        self.workspace = torch.empty(32 << 20, device=self.device, dtype=torch.uint8)
        L.check(self.lib.ivlm_set_workspace(self.h, P(self.workspace), C.c_size_t(self.workspace.numel()), self.stream),
                "set_workspace")

    def close(self):
        if getattr(self, "h", None):
            self.lib.ivlm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_option(self, name: str, value: int):
        L.check(self.lib.ivlm_set_option(self.h, name.encode(), i32(value)), "set_option")

    def launch_count(self) -> int:
        return int(self.lib.ivlm_launch_count(self.h))

    # ------------------------------------------------------------------ per-kernel timing (bench.py roofline pass)
    _PROFILED = ("gemm", "attention", "layernorm", "rmsnorm", "add_bcast", "silu_mul", "im2col_patch", "im2col_3x3",
                 "sam_relpos", "sam_attention", "attn_small", "embed_splice", "embed_gather", "gather_rows", "rope_kv_store", "decode_linear",
                 "decode_chain", "decode_attention", "fill_rows", "decode_prepare", "decode_finish", "argmax", "cam_gate", "upscale_hyper_dot", "bilinear", "finalize")

    def enable_profile(self):
        """Bracket every launch with CUDA events on the launching stream (no synchronisation until profile_report()).
        Adds a few microseconds of host time per launch, so bench.py uses it in a separate pass from the headline timing."""
        self._prof = []
        self.profiling = True
        for name in self._PROFILED:
            fn = getattr(type(self), name)

            def wrapped(*a, _fn=fn, _name=name, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = _fn(self, *a, **k)
                e1.record()
                work = 0.0
                if _name == "gemm":
                    work = 2.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0]
                    if a[0].shape[0] <= 64:  # swapped-operand weight streaming (decode, decoder tokens): HBM-bound
                        _name = "gemm_small_m"
                        work = 2.0 * a[1].shape[0] * a[1].shape[1]  # bytes of weights read
                elif _name == "decode_linear":
                    work = 2.0 * a[1].shape[0] * a[1].shape[1]  # bytes of weights streamed
                elif _name == "decode_chain":
                    work = sum(2.0 * ph[1].shape[0] * ph[1].shape[1] for ph in a[0])
                elif _name == "sam_attention":
                    Bq, nh, S_, hd_ = a[3], a[4], a[5] * a[6], a[7]
                    work = 4.0 * Bq * nh * S_ * S_ * hd_
                elif _name == "attention":
                    B, Sq, H, D = a[0].shape
                    work = 4.0 * B * H * Sq * a[1].shape[1] * D * (0.5 if k.get("causal") else 1.0)
                self._prof.append((_name, e0, e1, work))
                return r

            setattr(self, name, wrapped)

    def disable_profile(self):
        self.profiling = False
        for name in self._PROFILED:
            if name in self.__dict__:
                delattr(self, name)

    def profile_report(self) -> dict:
        """{kernel family: {"launches", "ms", "work" (FLOPs where defined)}} since enable_profile()."""
        torch.cuda.synchronize(self.device)
        rep = {}
        for name, e0, e1, work in self._prof:
            r = rep.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
            r["launches"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["work"] += work
        self._prof = []
        return rep

    # ------------------------------------------------------------------ stage-level ABI (csrc/stages.cu)
    def bind_weights(self, named: dict):
        """name -> CUDA tensor; pointers are borrowed (the caller keeps the tensors alive)."""
        arr = (L.WeightDesc * len(named))()
        keep = []
        for d, (name, t) in zip(arr, named.items()):
            assert t.is_cuda and t.is_contiguous() and t.dim() >= 1 and t.dim() <= 4, name
            nb = name.encode()
            keep.append(nb)
            d.name, d.ptr, d.ndim = nb, t.data_ptr(), t.dim()
            d.dtype = BF16 if t.dtype == torch.bfloat16 else (F32 if t.dtype == torch.float32 else 2)
            for k, sz in enumerate(t.shape):
                d.shape[k] = sz
        L.check(self.lib.ivlm_bind_weights(self.h, arr, i32(len(named))), "bind_weights")

    def set_model_dims(self, cfg, paired_layout: bool):
        d = L.ModelDims()
        d.sam_img, d.sam_patch, d.sam_embed_dim, d.sam_depth = cfg.sam_img_size, cfg.sam_patch_size, cfg.sam_embed_dim, cfg.sam_depth
        d.sam_heads, d.sam_window, d.sam_out_chans = cfg.sam_num_heads, cfg.sam_window_size, cfg.sam_out_chans
        d.sam_global_mask = sum(1 << i for i in cfg.sam_global_attn_indexes)
        d.llm_hidden, d.llm_intermediate, d.llm_layers = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        d.llm_heads, d.llm_head_dim, d.llm_vocab = cfg.num_attention_heads, cfg.head_dim, cfg.vocab_size
        d.llm_rms_eps, d.llm_paired_layout = cfg.rms_norm_eps, 1 if paired_layout else 0
        d.clip_img, d.clip_patch, d.clip_hidden = cfg.clip_image_size, cfg.clip_patch_size, cfg.clip_hidden_size
        d.clip_heads, d.clip_layers = cfg.clip_num_attention_heads, cfg.clip_layers_used
        d.clip_ldk, d.clip_eps = (3 * cfg.clip_patch_size ** 2 + 7) // 8 * 8, cfg.clip_layer_norm_eps
        L.check(self.lib.ivlm_set_model_dims(self.h, C.byref(d)), "set_model_dims")
        self._dims = d

    def sam_encode_stage(self, images, win_map, win_inv, win_pads):
        """ivlm_sam_encode: [N,3,S,S] bf16 -> [N, (S/patch)^2, out_chans] bf16 in one call."""
        _bf16(images)
        assert images.is_contiguous()
        N = images.shape[0]
        d = self._dims
        S = (d.sam_img // d.sam_patch) ** 2
        emb = torch.empty((N, S, d.sam_out_chans), device=images.device, dtype=torch.bfloat16)
        nbytes = int(self.lib.ivlm_sam_encode_arena_bytes(self.h, i32(N)))
        arena = torch.empty(nbytes, device=images.device, dtype=torch.uint8)
        a = L.SamEncodeArgs()
        a.images, a.emb, a.N = images.data_ptr(), emb.data_ptr(), N
        a.win_map, a.win_inv = win_map.data_ptr(), win_inv.data_ptr()
        a.win_pads, a.n_pads = (win_pads.data_ptr() if win_pads.numel() else None), win_pads.numel()
        a.arena, a.arena_bytes = arena.data_ptr(), nbytes
        L.check(self.lib.ivlm_sam_encode(self.h, C.byref(a), self.stream), "sam_encode")
        return emb

    def clip_encode_stage(self, images, patch_rows, cls_rows):
        """ivlm_clip_encode: [B,3,S,S] bf16 -> projected patch features [B, T-1, llm_hidden] bf16 in one call."""
        _bf16(images)
        assert images.is_contiguous() and patch_rows.dtype == torch.int32 and cls_rows.dtype == torch.int32
        B, d = images.shape[0], self._dims
        T1 = (d.clip_img // d.clip_patch) ** 2
        feats = torch.empty((B, T1, d.llm_hidden), device=images.device, dtype=torch.bfloat16)
        nbytes = int(self.lib.ivlm_clip_encode_arena_bytes(self.h, i32(B)))
        arena = torch.empty(nbytes, device=images.device, dtype=torch.uint8)
        a = L.ClipEncodeArgs()
        a.images, a.feats, a.B = images.data_ptr(), feats.data_ptr(), B
        a.patch_rows, a.cls_rows = patch_rows.data_ptr(), cls_rows.data_ptr()
        a.arena, a.arena_bytes = arena.data_ptr(), nbytes
        L.check(self.lib.ivlm_clip_encode(self.h, C.byref(a), self.stream), "clip_encode")
        return feats

    def seg_head_stage(self, hidden_rows, cam_params, V, out_dim=256):
        """ivlm_seg_head: [n, D] hidden rows (+ [n,V,5] camera parameters) -> prompt [n,V,out_dim], emb [n,out_dim]."""
        _bf16(hidden_rows)
        assert hidden_rows.is_contiguous() and (cam_params is None or cam_params.is_contiguous())
        n = hidden_rows.shape[0]
        prompt = torch.empty((n, V, out_dim), device=hidden_rows.device, dtype=torch.bfloat16)
        emb = torch.empty((n, out_dim), device=hidden_rows.device, dtype=torch.bfloat16)
        arena = torch.empty(n * hidden_rows.shape[1] * 2 + 4096, device=hidden_rows.device, dtype=torch.uint8)
        L.check(self.lib.ivlm_seg_head(self.h, P(hidden_rows), P(cam_params) if cam_params is not None else None, P(prompt), P(emb), i32(n), i32(V),
                                       P(arena), C.c_size_t(arena.numel()), self.stream), "seg_head")
        return prompt, emb

    def mask_decode_stage(self, emb, prompt, tok_idx, heads, grid):
        """ivlm_mask_decode: emb [n*V, S, C], prompt [n, V, C] -> low-res logits [n*V, 4*grid, 4*grid] fp32 in one call."""
        _bf16(emb); _bf16(prompt)
        assert emb.is_contiguous() and prompt.is_contiguous() and tok_idx.dtype == torch.int32
        n, V = prompt.shape[0], prompt.shape[1]
        low = torch.empty((n * V, 4 * grid, 4 * grid), device=emb.device, dtype=torch.float32)
        nbytes = int(self.lib.ivlm_mask_decode_arena_bytes(self.h, i32(n), i32(V)))
        arena = torch.empty(nbytes, device=emb.device, dtype=torch.uint8)
        a = L.MaskDecodeArgs()
        a.emb, a.prompt, a.lowres = emb.data_ptr(), prompt.data_ptr(), low.data_ptr()
        a.n, a.V, a.heads, a.tok_idx = n, V, heads, tok_idx.data_ptr()
        a.arena, a.arena_bytes = arena.data_ptr(), nbytes
        L.check(self.lib.ivlm_mask_decode(self.h, C.byref(a), self.stream), "mask_decode")
        return low

    def llm_arena(self, tokens, device):
        nbytes = int(self.lib.ivlm_llm_arena_bytes(self.h, i32(tokens)))
        return torch.empty(nbytes, device=device, dtype=torch.uint8)

    @staticmethod
    def pointer_array(tensors):
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return arr

    def llm_prefill_stage(self, embeds2d, positions, slot_map, k_ptrs, v_ptrs, hidden, next_tok, last_rows, B, S, page_size):
        _bf16(embeds2d)
        assert embeds2d.is_contiguous() and hidden.is_contiguous()
        arena = self.llm_arena(B * S, embeds2d.device)
        a = L.LlmPrefillArgs()
        a.embeds, a.positions, a.slot_map = embeds2d.data_ptr(), positions.data_ptr(), slot_map.data_ptr()
        a.k_cache, a.v_cache = k_ptrs, v_ptrs
        a.hidden, a.next_tok = hidden.data_ptr(), next_tok.data_ptr()
        a.last_rows = last_rows.data_ptr() if last_rows is not None else None
        a.B, a.S, a.max_len, a.page_size = B, S, hidden.shape[1], page_size
        a.arena, a.arena_bytes = arena.data_ptr(), arena.numel()
        L.check(self.lib.ivlm_llm_prefill(self.h, C.byref(a), self.stream), "llm_prefill")

    def llm_decode_stage(self, st, S, G, eos, pad, page_size):
        """ivlm_llm_decode_step over the decode-state dict of model.py (fixed buffers: CUDA-graph capturable)."""
        a = L.LlmDecodeArgs()
        a.state, a.S, a.G = st["state"].data_ptr(), S, G
        a.S_rows = st["S_rows"].data_ptr() if st.get("S_rows") is not None else None
        a.scripted = st["scripted"].data_ptr() if st.get("scripted") is not None else None
        for k in ("next", "done", "out_tokens", "tok", "pos", "slot", "seq_lens", "slot_base"):
            setattr(a, k, st[k].data_ptr())
        a.eos, a.pad, a.B = eos, pad, st["tok"].numel()
        a.k_cache, a.v_cache = st["k_ptrs"], st["v_ptrs"]
        a.block_table, a.max_pages, a.page_size = st["block_table"].data_ptr(), st["block_table"].shape[1], page_size
        a.hidden, a.max_len, a.hid_step = st["hidden"].data_ptr(), st["hidden"].shape[1], st["hid_step"].data_ptr()
        a.arena, a.arena_bytes = st["decode_arena"].data_ptr(), st["decode_arena"].numel()
        L.check(self.lib.ivlm_llm_decode_step(self.h, C.byref(a), self.stream), "llm_decode_step")

    # ------------------------------------------------------------------ dense
    def gemm(self, a, w, bias=None, act=ACT_NONE, residual=None, out=None, out_dtype=torch.bfloat16, row_map=None,
             out_rows=None, k_splits=0, force_swap=0, no_round=False, res_row_mod=0):
        """out = act(a @ w.T + bias) + residual.  a [M,K], w [N,K] bf16 (last dim contiguous)."""
        _bf16(a, "a"); _bf16(w, "w")
        assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
        assert a.stride(1) == 1 and w.stride(1) == 1
        M, K = a.shape
        N = w.shape[0]
        if out is None:
            rows = M if out_rows is None else out_rows
            out = torch.empty((rows, N), device=a.device, dtype=out_dtype)
            if k_splits > 1:
                out.zero_()
        assert out.stride(1) == 1
        args = L.GemmArgs()
        args.a, args.lda = a.data_ptr(), a.stride(0)
        args.w, args.ldw = w.data_ptr(), w.stride(0)
        args.out, args.ldo = out.data_ptr(), out.stride(0)
        args.bias = _bf16(bias, "bias").data_ptr() if bias is not None else None
        if residual is not None:
            _bf16(residual, "residual")
            assert residual.stride(1) == 1
            args.residual, args.ldr = residual.data_ptr(), residual.stride(0)
        if row_map is not None:
            assert row_map.dtype == torch.int32 and row_map.numel() == M
            args.row_map = row_map.data_ptr()
        args.M, args.N, args.K = M, N, K
        args.act = act
        args.out_dtype = BF16 if out.dtype == torch.bfloat16 else F32
        args.k_splits = k_splits
        args.force_swap = force_swap
        args.no_round = 1 if no_round else 0
        args.res_row_mod = res_row_mod
        L.check(self.lib.ivlm_gemm_bf16(self.h, C.byref(args), self.stream), "gemm")
        return out

    # ------------------------------------------------------------------ norms / elementwise
    def layernorm(self, x, gamma, beta, eps, row_map=None, out_rows=None, act=ACT_NONE, out=None):
        _bf16(x)
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        rows = x2.shape[0] if row_map is None else (row_map.numel() if out_rows is None else out_rows)
        if out is None:
            out = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_layernorm_bf16(self.h, P(x2), P(out), P(gamma), P(beta), i64(rows), i32(D), f32c(eps),
                                             P(row_map), i32(act), self.stream), "layernorm")
        return out if row_map is not None else out.view(x.shape)

    def rmsnorm(self, x, gamma, eps, out=None):
        _bf16(x)
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        if out is None:
            out = torch.empty_like(x2)
        L.check(self.lib.ivlm_rmsnorm_bf16(self.h, P(x2), P(out), P(gamma), i64(x2.shape[0]), i32(D), f32c(eps),
                                           self.stream), "rmsnorm")
        return out.view(x.shape)

    def add_bcast(self, a, b, out=None):
        _bf16(a); _bf16(b)
        assert a.is_contiguous() and b.is_contiguous()
        if out is None:
            out = torch.empty_like(a)
        period = 0 if b.numel() == a.numel() else b.numel()
        L.check(self.lib.ivlm_add_bcast_bf16(self.h, P(a), P(b), P(out), i64(a.numel()), i64(period), self.stream),
                "add_bcast")
        return out

    def silu_mul(self, gate_up, out=None, interleaved=False):
        _bf16(gate_up)
        rows, F2 = gate_up.shape
        if out is None:
            out = torch.empty((rows, F2 // 2), device=gate_up.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_silu_mul_bf16(self.h, P(gate_up), P(out), i64(rows), i32(F2 // 2), i32(1 if interleaved else 0),
                                            self.stream), "silu_mul")
        return out

    def _decode_linear_args(self, a, w, gamma=None, eps=0.0, epilogue=L.EPI_PLAIN, act=ACT_NONE, bias=None, residual=None, out=None,
                            out_dtype=torch.bfloat16, rope=None, prefetch=None, prefetch_stages=0):
        _bf16(a, "a"); _bf16(w, "w")
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
        width = N // 2 if epilogue == L.EPI_SWIGLU else (N // 3 if epilogue == L.EPI_ROPE_KV else N)
        if out is None:
            out = torch.empty((M, width), device=a.device, dtype=out_dtype)
        assert out.shape == (M, width) and out.stride(1) == 1
        g = L.DecodeLinearArgs()
        g.a, g.lda, g.w, g.ldw = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
        g.M, g.N, g.K = M, N, K
        g.norm_gamma, g.norm_eps = (_bf16(gamma, "gamma").data_ptr() if gamma is not None else None), float(eps)
        g.epilogue, g.act = epilogue, act
        g.bias = _bf16(bias, "bias").data_ptr() if bias is not None else None
        if residual is not None:
            _bf16(residual, "residual")
            g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
        g.out, g.ldo = out.data_ptr(), out.stride(0)
        g.out_dtype = BF16 if out.dtype == torch.bfloat16 else F32
        if epilogue == L.EPI_ROPE_KV:
            assert rope["positions"].dtype == torch.int32 and rope["slot_map"].dtype == torch.int32
            g.positions, g.slot_map = rope["positions"].data_ptr(), rope["slot_map"].data_ptr()
            g.cos_t, g.sin_t = rope["cos"].data_ptr(), rope["sin"].data_ptr()
            g.k_cache, g.v_cache = rope["k_cache"].data_ptr(), rope["v_cache"].data_ptr()
            g.H, g.hd, g.page_size = rope["H"], rope["hd"], rope["page_size"]
        if prefetch is not None:
            _bf16(prefetch, "prefetch")
            assert prefetch.dim() == 2 and prefetch.stride(1) == 1
            g.prefetch_w, g.prefetch_ldw = prefetch.data_ptr(), prefetch.stride(0)
            g.prefetch_N, g.prefetch_K, g.prefetch_stages = prefetch.shape[0], prefetch.shape[1], int(prefetch_stages)
        return g, out

    def decode_linear(self, a, w, **kw):
        """Weight-streaming linear layer of a decode step, M <= 8 tokens (ivlm_decode_linear): optional fused RMSNorm of `a`
        (gamma, eps), epilogue PLAIN (act, bias, residual) / SWIGLU (w rows interleaved) / ROPE_KV (w q,k rows paired; rope =
        dict(positions, slot_map, cos, sin, k_cache, v_cache, H, hd, page_size)); out / out_dtype.  `prefetch` = the weight matrix
        of the next decode_linear launch (its first stages are requested into L2 at the end of this one when the option
        ds_prefetch_kb enables it; prefetch_stages 16 KB stages per SM, 0 = library default)."""
        g, out = self._decode_linear_args(a, w, **kw)
        L.check(self.lib.ivlm_decode_linear(self.h, C.byref(g), self.stream), "decode_linear")
        return out

    def decode_chain(self, phases):
        """Up to four dependent decode_linear launches as one persistent launch with grid-wide barriers between the phases
        (ivlm_decode_chain).  phases: list of (a, w, kwargs) with decode_linear's keywords; a phase reads the `out` tensor of an
        earlier one by passing it as `a` / `residual` (allocate the outputs first and pass them as out=).  Returns the outputs."""
        arr = (L.DecodeLinearArgs * len(phases))()
        outs = []
        for i, (a, w, kw) in enumerate(phases):
            g, out = self._decode_linear_args(a, w, **kw)
            arr[i] = g
            outs.append(out)
        L.check(self.lib.ivlm_decode_chain(self.h, arr, i32(len(phases)), self.stream), "decode_chain")
        return outs

    def finalize(self, acc, bias=None, residual=None, act=ACT_NONE, out=None):
        assert acc.dtype == torch.float32 and acc.is_contiguous()
        rows, N = acc.shape
        if out is None:
            out = torch.empty((rows, N), device=acc.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_finalize_f32_bf16(self.h, P(acc), P(out), P(bias), P(residual), i64(rows), i32(N), i32(act),
                                                self.stream), "finalize")
        return out

    def to_bf16(self, x):
        assert x.dtype == torch.float32 and x.is_contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_cast_f32_bf16(self.h, P(x), P(out), i64(x.numel()), self.stream), "cast")
        return out

    def to_f32(self, x):
        _bf16(x)
        assert x.is_contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        L.check(self.lib.ivlm_cast_bf16_f32(self.h, P(x), P(out), i64(x.numel()), self.stream), "cast")
        return out

    # ------------------------------------------------------------------ lowering
    def im2col_patch(self, img, p, ldk=None):
        _bf16(img)
        assert img.is_contiguous()
        N, Cc, H, W = img.shape
        kk = Cc * p * p
        ldk = ldk or ((kk + 7) // 8) * 8
        out = torch.empty((N * (H // p) * (W // p), ldk), device=img.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_im2col_patch_bf16(self.h, P(img), P(out), i32(N), i32(Cc), i32(H), i32(W), i32(p), i32(ldk),
                                                self.stream), "im2col_patch")
        return out

    def im2col_3x3(self, x, N, H, W):
        _bf16(x)
        assert x.is_contiguous()
        Cc = x.shape[-1]
        out = torch.empty((N * H * W, 9 * Cc), device=x.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_im2col_3x3_bf16(self.h, P(x), P(out), i32(N), i32(H), i32(W), i32(Cc), self.stream),
                "im2col_3x3")
        return out

    # ------------------------------------------------------------------ attention
    def attention(self, q, k, v, scale, causal=False, rel_h=None, rel_w=None, kh=0, kw=0, out=None):
        """q [B,Sq,H,D], k/v [B,Sk,H,D] bf16 views (any batch/token/head strides, D contiguous) -> [B,Sq,H,D]."""
        _bf16(q); _bf16(k); _bf16(v)
        B, Sq, H, D = q.shape
        Sk = k.shape[1]
        assert q.stride(3) == 1 and k.stride(3) == 1 and v.stride(3) == 1
        if out is None:
            out = torch.empty((B, Sq, H, D), device=q.device, dtype=torch.bfloat16)
        a = L.AttnArgs()
        a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
        a.q_bs, a.q_ts, a.q_hs = q.stride(0), q.stride(1), q.stride(2)
        a.k_bs, a.k_ts, a.k_hs = k.stride(0), k.stride(1), k.stride(2)
        a.v_bs, a.v_ts, a.v_hs = v.stride(0), v.stride(1), v.stride(2)
        a.o_bs, a.o_ts, a.o_hs = out.stride(0), out.stride(1), out.stride(2)
        a.B, a.H, a.Sq, a.Sk, a.D = B, H, Sq, Sk, D
        a.scale = scale
        a.causal = 1 if causal else 0
        if rel_h is not None:
            assert rel_h.dtype == torch.float32 and rel_w.dtype == torch.float32
            a.rel_h, a.rel_w, a.kh, a.kw = rel_h.data_ptr(), rel_w.data_ptr(), kh, kw
        L.check(self.lib.ivlm_attention_bf16(self.h, C.byref(a), self.stream), "attention")
        return out

    def sam_relpos(self, qkv, rel_pos_h, rel_pos_w, B, heads, Hq, Wq, hd):
        _bf16(qkv)
        S = Hq * Wq
        rel_h = torch.empty((B, heads, S, Hq), device=qkv.device, dtype=torch.float32)
        rel_w = torch.empty((B, heads, S, Wq), device=qkv.device, dtype=torch.float32)
        L.check(self.lib.ivlm_sam_relpos(self.h, P(qkv), P(rel_pos_h), P(rel_pos_w), P(rel_h), P(rel_w), i32(B),
                                         i32(heads), i32(Hq), i32(Wq), i32(hd), self.stream), "sam_relpos")
        return rel_h, rel_w

    def sam_attention(self, qkv, rel_pos_h, rel_pos_w, B, heads, Hq, Wq, hd, out=None, out_map=None, out_rows=None):
        """Fused SAM attention (tcgen05): qkv [B*Hq*Wq, 3*heads*hd] -> [B*Hq*Wq, heads*hd]; with out_map ([B*Hq*Wq] int32,
        windows only) row r is stored at out_map[r] of an [out_rows, heads*hd] output (-1: dropped)."""
        _bf16(qkv); _bf16(rel_pos_h); _bf16(rel_pos_w)
        assert qkv.is_contiguous() and rel_pos_h.is_contiguous() and rel_pos_w.is_contiguous()
        assert qkv.shape == (B * Hq * Wq, 3 * heads * hd), (qkv.shape, B, Hq, Wq, heads, hd)
        if out is None:
            rows = B * Hq * Wq if out_map is None else out_rows
            out = torch.empty((rows, heads * hd), device=qkv.device, dtype=torch.bfloat16)
        if out_map is not None:
            assert out_map.dtype == torch.int32 and out_map.numel() == B * Hq * Wq
        L.check(self.lib.ivlm_sam_attention_bf16(self.h, P(qkv), P(rel_pos_h), P(rel_pos_w), P(out), i32(B), i32(heads),
                                                 i32(Hq), i32(Wq), i32(hd), i64(out.stride(0)), P(out_map), self.stream),
                "sam_attention")
        return out

    def fill_rows(self, out, rows, vec):
        """out[rows[i], :] = vec (bf16) -- bias broadcast into the rows of the SAM window padding."""
        _bf16(out); _bf16(vec)
        assert rows.dtype == torch.int32 and out.stride(1) == 1 and vec.numel() == out.shape[1]
        L.check(self.lib.ivlm_fill_rows_bf16(self.h, P(out), i64(out.stride(0)), P(rows), i32(rows.numel()), P(vec),
                                             i32(out.shape[1]), self.stream), "fill_rows")
        return out

    def attn_small(self, q, k, v, heads):
        """q [Bq,Nq,C] (Bq == 1 broadcasts), k/v [B,Nk,C] -> [B,Nq,C]."""
        _bf16(q); _bf16(k); _bf16(v)
        assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
        B, Nk, Cc = k.shape
        Nq = q.shape[1]
        bcast = 1 if (q.shape[0] == 1 and B > 1) else 0
        out = torch.empty((B, Nq, Cc), device=q.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_attn_small_bf16(self.h, P(q), P(k), P(v), P(out), i32(B), i32(bcast), i32(Nq), i32(Nk),
                                              i32(heads), i32(Cc // heads), self.stream), "attn_small")
        return out

    # ------------------------------------------------------------------ LLaVA / LLaMA glue
    def embed_splice(self, embed, ids, img_feats):
        _bf16(embed); _bf16(img_feats)
        assert ids.dtype == torch.int32 and ids.is_contiguous() and img_feats.is_contiguous()
        B, Lq = ids.shape
        n_img, D = img_feats.shape[1], img_feats.shape[2]
        out = torch.empty((B, Lq - 1 + n_img, D), device=embed.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_embed_splice_bf16(self.h, P(embed), P(ids), P(img_feats), P(out), i32(B), i32(Lq), i32(n_img),
                                                i32(D), i32(embed.shape[0]), self.stream), "embed_splice")
        return out

    def embed_gather(self, embed, ids, out=None):
        assert ids.dtype == torch.int32
        n, D = ids.numel(), embed.shape[1]
        if out is None:
            out = torch.empty((n, D), device=embed.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_embed_gather_bf16(self.h, P(embed), P(ids), P(out), i32(n), i32(D), i32(embed.shape[0]),
                                                self.stream), "embed_gather")
        return out

    def rows_differ(self, x, ref):
        """x [n, ...], ref [K, ...] (same row shape and dtype, contiguous) -> int32 [n, K], 1 where the rows differ bitwise."""
        assert x.is_cuda and ref.is_cuda and x.dtype == ref.dtype and x.shape[1:] == ref.shape[1:]
        assert x.is_contiguous() and ref.is_contiguous()
        n, K = x.shape[0], ref.shape[0]
        row_bytes = x[0].numel() * x.element_size()
        neq = torch.empty((n, K), device=x.device, dtype=torch.int32)
        L.check(self.lib.ivlm_rows_differ(self.h, P(x), i32(n), P(ref), i32(K), i64(row_bytes), P(neq), self.stream), "rows_differ")
        return neq

    def gather_rows(self, x, idx):
        _bf16(x)
        assert idx.dtype == torch.int32 and x.is_contiguous()
        out = torch.empty((idx.numel(), x.shape[-1]), device=x.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_gather_rows_bf16(self.h, P(x), P(idx), P(out), i32(idx.numel()), i32(x.shape[-1]),
                                               self.stream), "gather_rows")
        return out

    def rope_kv_store(self, qkv, positions, slot_map, cos_t, sin_t, H, hd, k_cache=None, v_cache=None, want_kv=True,
                      q_out=None, page_size=16, paired=False):
        _bf16(qkv)
        T = qkv.shape[0]
        D = H * hd
        if q_out is None:
            q_out = torch.empty((T, D), device=qkv.device, dtype=torch.bfloat16)
        k_out = torch.empty((T, D), device=qkv.device, dtype=torch.bfloat16) if want_kv else None
        v_out = torch.empty((T, D), device=qkv.device, dtype=torch.bfloat16) if want_kv else None
        L.check(self.lib.ivlm_rope_kv_store_bf16(self.h, P(qkv), P(positions), P(slot_map), P(cos_t), P(sin_t), P(q_out),
                                                 P(k_out), P(v_out), P(k_cache), P(v_cache), i32(T), i32(H), i32(hd),
                                                 i32(page_size), i32(1 if paired else 0), self.stream), "rope_kv_store")
        return q_out, k_out, v_out

    def decode_attention(self, q, k_cache, v_cache, block_table, seq_lens, H, hd, page_size, out=None):
        _bf16(q)
        B = q.shape[0]
        if out is None:
            out = torch.empty((B, H * hd), device=q.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_decode_attention_paged_bf16(self.h, P(q), P(k_cache), P(v_cache), P(block_table), P(seq_lens),
                                                          P(out), i32(B), i32(H), i32(hd), i32(page_size),
                                                          i32(block_table.shape[1]), f32c(1.0 / math.sqrt(hd)),
                                                          self.stream), "decode_attention")
        return out

    def decode_prepare(self, st, S, G, eos, pad):
        """Device-side token / position bookkeeping of one decode step (st: the model's decode-state dict)."""
        B = st["tok"].numel()
        L.check(self.lib.ivlm_decode_prepare(self.h, P(st["state"]), i32(S), P(st.get("S_rows")), P(st.get("scripted")), i32(G), P(st["next"]),
                                             P(st["done"]), P(st["out_tokens"]), P(st["tok"]), P(st["pos"]), P(st["slot"]),
                                             P(st["seq_lens"]), P(st["slot_base"]), i32(eos), i32(pad), i32(B), self.stream),
                "decode_prepare")

    def decode_finish(self, st, S):
        hid = st["hidden"]
        L.check(self.lib.ivlm_decode_finish(self.h, P(st["state"]), i32(S), P(st.get("S_rows")), P(st["hid_step"]), P(hid), i32(hid.shape[0]),
                                            i32(hid.shape[2]), i32(hid.shape[1]), self.stream), "decode_finish")

    def argmax(self, logits, vocab=None, out=None):
        assert logits.dtype == torch.float32 and logits.stride(1) == 1
        B = logits.shape[0]
        vocab = vocab or logits.shape[1]
        if out is None:
            out = torch.empty((B,), device=logits.device, dtype=torch.int32)
        L.check(self.lib.ivlm_argmax_f32(self.h, P(logits), P(out), i32(B), i32(vocab), i64(logits.stride(0)),
                                         self.stream), "argmax")
        return out

    # ------------------------------------------------------------------ prompt / mask tail
    def cam_gate(self, cam, emb, w1, b1, w2, b2, wv, bv):
        _bf16(cam); _bf16(emb)
        B, V = cam.shape[0], cam.shape[1]
        out = torch.empty((B, V, 256), device=cam.device, dtype=torch.bfloat16)
        L.check(self.lib.ivlm_cam_gate_bf16(self.h, P(cam.contiguous()), P(emb.contiguous()), P(w1), P(b1), P(w2), P(b2),
                                            P(wv), P(bv), P(out), i32(B), i32(V), self.stream), "cam_gate")
        return out

    def upscale_hyper_dot(self, up1, w2, b2, hyper, Bv, grid):
        _bf16(up1); _bf16(hyper)
        out = torch.empty((Bv, grid * 4, grid * 4), device=up1.device, dtype=torch.float32)
        L.check(self.lib.ivlm_upscale_hyper_dot(self.h, P(up1), P(w2), P(b2), P(hyper.contiguous()), P(out), i32(Bv),
                                                i32(grid), self.stream), "upscale_hyper_dot")
        return out

    SAM_MEAN, SAM_STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
    CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)

    def preprocess_u8(self, img_u8, size, kind="sam"):
        """uint8 [N,H,W,3] device tensor -> normalised bf16 [N,3,size,size] (zero padded).  kind: 'sam' | 'clip'."""
        assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.is_contiguous() and img_u8.shape[-1] == 3
        N, H, W, _ = img_u8.shape
        mean, std, pre = (self.SAM_MEAN, self.SAM_STD, 1.0) if kind == "sam" else (self.CLIP_MEAN, self.CLIP_STD, 1.0 / 255.0)
        out = torch.empty((N, 3, size, size), device=img_u8.device, dtype=torch.bfloat16)
        m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        L.check(self.lib.ivlm_preprocess_u8_bf16(self.h, P(img_u8), P(out), i32(N), i32(H), i32(W), i32(size), f32c(pre), m3, s3,
                                                 self.stream), "preprocess_u8")
        return out

    def decode_jpeg(self, data: bytes):
        """Compressed JPEG bytes -> uint8 [H,W,3] RGB CUDA tensor (nvJPEG)."""
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        H, W = i32(0), i32(0)
        L.check(self.lib.ivlm_jpeg_info(self.h, buf, C.c_size_t(len(data)), C.byref(H), C.byref(W)), "jpeg_info")
        out = torch.empty((H.value, W.value, 3), device=self.device, dtype=torch.uint8)
        L.check(self.lib.ivlm_jpeg_decode_rgb(self.h, buf, C.c_size_t(len(data)), P(out), H, W, self.stream), "jpeg_decode")
        torch.cuda.current_stream(self.device).synchronize()   # `buf` (host) must outlive the decode
        return out

    def resize_u8(self, img_u8, out_h, out_w, filt="bilinear"):
        """Pillow-exact antialiased resize of uint8 [N,H,W,3] device images (horizontal pass, then vertical)."""
        from . import resample as R

        assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.is_contiguous() and img_u8.shape[-1] == 3
        N, H, W, _ = img_u8.shape
        cache = self.__dict__.setdefault("_resample_tables", {})

        def tables(n_in, n_out):
            key = (n_in, n_out, filt)
            if key not in cache:
                b, k, ks = R.precompute_coeffs(n_in, n_out, filt)
                cache[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), ks)
            return cache[key]

        x = img_u8
        if out_w != W:
            b, k, ks = tables(W, out_w)
            y = torch.empty((N, H, out_w, 3), device=x.device, dtype=torch.uint8)
            L.check(self.lib.ivlm_resample_u8(self.h, P(x), P(y), P(b), P(k), i32(ks), i32(N), i32(H), i32(W), i32(H), i32(out_w),
                                              i32(0), self.stream), "resample_u8")
            x = y
        if out_h != H:
            b, k, ks = tables(H, out_h)
            y = torch.empty((N, out_h, out_w, 3), device=x.device, dtype=torch.uint8)
            L.check(self.lib.ivlm_resample_u8(self.h, P(x), P(y), P(b), P(k), i32(ks), i32(N), i32(H), i32(out_w), i32(out_h),
                                              i32(out_w), i32(1), self.stream), "resample_u8")
            x = y
        return x

    def sigmoid_where(self, x, gt=None, ignore_value=-1.0):
        """In place: x = sigmoid(x) where gt != ignore_value (everywhere when gt is None)."""
        assert x.dtype == torch.float32 and x.is_contiguous()
        if gt is not None:
            assert gt.dtype == torch.float32 and gt.is_contiguous() and gt.numel() == x.numel()
        L.check(self.lib.ivlm_sigmoid_where_f32(self.h, P(x), P(gt), f32c(ignore_value), i64(x.numel()), self.stream),
                "sigmoid_where")
        return x

    def bilinear(self, src, dh, dw, crop_h=None, crop_w=None, out=None):
        assert src.dtype == torch.float32 and src.is_contiguous() and src.dim() == 3
        N, sh, sw = src.shape
        crop_h, crop_w = crop_h or sh, crop_w or sw
        if out is None:
            out = torch.empty((N, dh, dw), device=src.device, dtype=torch.float32)
        L.check(self.lib.ivlm_bilinear_f32(self.h, P(src), P(out), i32(N), i32(sh), i32(sw), i32(crop_h), i32(crop_w),
                                           i32(dh), i32(dw), self.stream), "bilinear")
        return out


def _cam_array(cams):
    """list of dict(R [3,3], T [3], C [3], s | (fx, fy, cx, cy), z_clip) -> ctypes array of ivlm_raster_cam."""
    assert 1 <= len(cams) <= L.RASTER_MAX_VIEWS, f"1..{L.RASTER_MAX_VIEWS} cameras per call"
    arr = (L.RasterCam * len(cams))()
    for a, c in zip(arr, cams):
        a.R[:] = [float(x) for x in np.asarray(c["R"], dtype=np.float32).reshape(9)]
        a.T[:] = [float(x) for x in np.asarray(c["T"], dtype=np.float32).reshape(3)]
        a.C[:] = [float(x) for x in np.asarray(c["C"], dtype=np.float32).reshape(3)]
        a.fx, a.fy = float(c.get("fx", c.get("s", 1.0))), float(c.get("fy", c.get("s", 1.0)))
        a.cx, a.cy, a.z_clip = float(c.get("cx", 0.0)), float(c.get("cy", 0.0)), float(c["z_clip"])
    return arr


def rasterize_mesh(ctx: Context, verts, faces, cams, H, W, want_p2v=True, want_zbuf=False):
    """verts [Nv,3] fp32, faces [Nf,3] int32 (CUDA) -> dict(pix_to_face [V,H,W] i32, bary [V,H,W,3] f32,
    p2v [V,H,W,3] i64 | None, zbuf | None, skipped).  pytorch3d MeshRasterizer semantics (include/ivlm_b200.h)."""
    assert verts.is_cuda and verts.dtype == torch.float32 and verts.is_contiguous() and verts.shape[1] == 3
    assert faces.is_cuda and faces.dtype == torch.int32 and faces.is_contiguous() and faces.shape[1] == 3
    V, dev = len(cams), verts.device
    pix = torch.empty((V, H, W), device=dev, dtype=torch.int32)
    bary = torch.empty((V, H, W, 3), device=dev, dtype=torch.float32)
    zbuf = torch.empty((V, H, W), device=dev, dtype=torch.float32) if want_zbuf else None
    p2v = torch.empty((V, H, W, 3), device=dev, dtype=torch.int64) if want_p2v else None
    skipped = i32(0)
    L.check(ctx.lib.ivlm_rasterize_mesh(ctx.h, P(verts), P(faces), i32(verts.shape[0]), i32(faces.shape[0]), _cam_array(cams),
                                        i32(V), i32(H), i32(W), P(pix), P(bary), P(zbuf), P(p2v), C.byref(skipped), ctx.stream),
            "rasterize_mesh")
    return dict(pix_to_face=pix, bary=bary, zbuf=zbuf, p2v=p2v, skipped=int(skipped.value))


def rasterize_points(ctx: Context, points, cams, H, W, radius):
    """points [n,3] fp32 CUDA -> pixel_to_point [V,H,W] int64 (-1 background): nearest point in depth within `radius` (NDC) of
    each pixel centre (ivlm_rasterize_points)."""
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2 and points.shape[1] == 3
    V = len(cams)
    p2p = torch.empty((V, H, W), device=points.device, dtype=torch.int64)
    L.check(ctx.lib.ivlm_rasterize_points(ctx.h, P(points), i32(points.shape[0]), _cam_array(cams), i32(V), i32(H), i32(W),
                                          f32c(float(radius)), P(p2p), ctx.stream), "rasterize_points")
    return p2p


def shade_phong(ctx: Context, verts, faces, colors, cams, lights, pix_to_face, bary, ambient=0.5, diffuse=0.3, specular=0.2,
                shininess=64.0):
    """HardPhongShader over the rasteriser output -> uint8 [V,H,W,3] (white background)."""
    assert colors.is_cuda and colors.dtype == torch.float32 and colors.is_contiguous() and colors.shape == verts.shape
    V, H, W = pix_to_face.shape
    lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(V, 3)
    rgb = torch.empty((V, H, W, 3), device=verts.device, dtype=torch.uint8)
    L.check(ctx.lib.ivlm_shade_phong(ctx.h, P(verts), P(faces), i32(verts.shape[0]), i32(faces.shape[0]), P(colors),
                                     _cam_array(cams), lights.ctypes.data_as(C.c_void_p), i32(V), i32(H), i32(W),
                                     P(pix_to_face), P(bary), f32c(ambient), f32c(diffuse), f32c(specular), f32c(shininess),
                                     P(rgb), ctx.stream), "shade_phong")
    return rgb


class LiftMap:
    """Per-(view, vertex) CSR built once from the reference's pixel->vertex / barycentric maps."""

    def __init__(self, ctx: Context, p2v: np.ndarray, bary: np.ndarray | None, n_verts: int):
        self.ctx = ctx
        p2v = np.ascontiguousarray(p2v, dtype=np.int64)
        self.ptr = C.c_void_p()
        if bary is not None:
            bary = np.ascontiguousarray(bary, dtype=np.float32)
            V, H, W, three = p2v.shape
            assert three == 3 and bary.shape == p2v.shape
            L.check(ctx.lib.ivlm_lift_build_mesh(ctx.h, p2v.ctypes.data_as(C.c_void_p), bary.ctypes.data_as(C.c_void_p),
                                                 i32(V), i32(H), i32(W), i32(n_verts), C.byref(self.ptr)),
                    "lift_build_mesh")
        else:
            V, H, W = p2v.shape
            L.check(ctx.lib.ivlm_lift_build_points(ctx.h, p2v.ctypes.data_as(C.c_void_p), i32(V), i32(H), i32(W),
                                                   i32(n_verts), C.byref(self.ptr)), "lift_build_points")
        self.V, self.H, self.W, self.n = V, H, W, n_verts
        self.nnz = int(ctx.lib.ivlm_lift_nnz(self.ptr))

    def __call__(self, masks: torch.Tensor, mode: int, thr: float = 0.3) -> torch.Tensor:
        assert masks.dtype == torch.float32 and masks.is_cuda and masks.is_contiguous()
        B = masks.shape[0]
        assert tuple(masks.shape[1:]) == (self.V, self.H, self.W), (masks.shape, (self.V, self.H, self.W))
        out = torch.empty((B, self.n), device=masks.device, dtype=torch.float32)
        L.check(self.ctx.lib.ivlm_lift(self.ctx.h, self.ptr, P(masks), P(out), i32(B), i32(mode), f32c(thr),
                                       self.ctx.stream), "lift")
        return out

    def lowres(self, low: torch.Tensor, mode: int, thr: float = 0.3) -> torch.Tensor:
        """low [B,V,sh,sw] fp32 low-res logits -> [B,n]: bilinear to (H,W) fused into the gather (ivlm_lift_lowres)."""
        assert low.dtype == torch.float32 and low.is_cuda and low.is_contiguous() and low.dim() == 4
        B = low.shape[0]
        assert low.shape[1] == self.V, (low.shape, self.V)
        out = torch.empty((B, self.n), device=low.device, dtype=torch.float32)
        L.check(self.ctx.lib.ivlm_lift_lowres(self.ctx.h, self.ptr, P(low), i32(low.shape[2]), i32(low.shape[3]), P(out), i32(B),
                                              i32(mode), f32c(thr), self.ctx.stream), "lift_lowres")
        return out

    def __del__(self):
        try:
            if self.ptr:
                self.ctx.lib.ivlm_lift_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class CsrMatrix:
    """Sparse form of a dense host matrix (SMPL -> SMPL-X vertex mapping, utils/utils.py:428-443)."""

    def __init__(self, ctx: Context, dense: np.ndarray):
        self.ctx = ctx
        dense = np.ascontiguousarray(dense, dtype=np.float32)
        self.rows, self.cols = dense.shape
        self.ptr = C.c_void_p()
        L.check(ctx.lib.ivlm_csr_build_dense(ctx.h, dense.ctypes.data_as(C.c_void_p), i32(self.rows), i32(self.cols),
                                             C.byref(self.ptr)), "csr_build_dense")

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous() and x.shape[-1] == self.cols
        B = x.reshape(-1, self.cols).shape[0]
        y = torch.empty((B, self.rows), device=x.device, dtype=torch.float32)
        L.check(self.ctx.lib.ivlm_csr_spmv(self.ctx.h, self.ptr, P(x), P(y), i32(B), self.ctx.stream), "csr_spmv")
        return y

    def __del__(self):
        try:
            if self.ptr:
                self.ctx.lib.ivlm_csr_free(self.ptr)
                self.ptr = None
        except Exception:
            pass
