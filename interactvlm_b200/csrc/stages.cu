// Stage-level entry points of the C ABI: a reference maintainer binds the checkpoint tensors once and then
// calls one function per stage -- the SAM ViT encoder, the LLaMA prefill, one LLaMA decode step -- instead of driving the
// 32-block / 40-layer loops op by op.  Each driver is the fixed launch sequence of the op-level entry points of this library
// (the same kernels, the same order as interactvlm_b200/model.py's _Engine: results are bit-identical to the op-level path);
// activations live in a caller-provided arena (bump-allocated, nothing is cudaMalloc'ed), all launches go to `stream`.
#include <string.h>

#include <string>

#include "runtime.h"

namespace ivlm {

struct Arena {
    char* base;
    size_t bytes, off = 0;
    Arena(void* p, size_t n) : base(reinterpret_cast<char*>(p)), bytes(n) {}
    void* take(size_t n) {
        const size_t a = (off + 255) & ~size_t(255);
        if (a + n > bytes) return nullptr;
        off = a + n;
        return base + a;
    }
};

static const Weight* find_weight(ivlm_ctx* h, const std::string& name) {
    auto it = h->weights.find(name);
    return it == h->weights.end() ? nullptr : &it->second;
}

#define IVLM_W(var, name)                                                                          \
    const Weight* var##_w = find_weight(h, name);                                                  \
    IVLM_REQUIRE(var##_w != nullptr, "stage driver: weight '%s' is not bound (ivlm_bind_weights)", std::string(name).c_str()); \
    const void* var = var##_w->ptr

#define IVLM_TAKE(var, type, count)                                                                \
    type* var = reinterpret_cast<type*>(arena.take(sizeof(type) * (size_t)(count)));               \
    IVLM_REQUIRE(var != nullptr, "stage driver: arena too small (%zu bytes given)", arena.bytes)

static int gemm(ivlm_ctx* h, const void* a, int64_t lda, const void* w, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                const void* bias, int act, const void* res, int64_t ldr, const int32_t* row_map, int res_row_mod, int force_swap,
                int out_dtype, void* stream) {
    ivlm_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.a = a; g.lda = lda; g.w = w; g.ldw = ldw; g.out = out; g.ldo = ldo;
    g.bias = bias; g.residual = res; g.ldr = ldr; g.row_map = row_map;
    g.M = M; g.N = N; g.K = K; g.act = act; g.out_dtype = out_dtype; g.force_swap = force_swap; g.res_row_mod = res_row_mod;
    return ivlm_gemm_bf16(h, &g, stream);
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_bind_weights(ivlm_handle h, const ivlm_weight_desc* descs, int32_t n) {
    IVLM_REQUIRE(h && (descs || n == 0) && n >= 0, "bind_weights: bad arguments");
    for (int i = 0; i < n; ++i) {
        IVLM_REQUIRE(descs[i].name && descs[i].ptr && descs[i].ndim >= 1 && descs[i].ndim <= 4, "bind_weights: entry %d is malformed", i);
        Weight w;
        w.ptr = descs[i].ptr;
        w.dtype = descs[i].dtype;
        w.ndim = descs[i].ndim;
        for (int d = 0; d < descs[i].ndim; ++d) w.shape[d] = descs[i].shape[d];
        h->weights[descs[i].name] = w;
    }
    return IVLM_OK;
}

extern "C" int ivlm_set_model_dims(ivlm_handle h, const ivlm_model_dims* d) {
    IVLM_REQUIRE(h && d, "set_model_dims: null");
    IVLM_REQUIRE(d->sam_embed_dim % d->sam_heads == 0 && d->sam_img % d->sam_patch == 0 && d->llm_hidden == d->llm_heads * d->llm_head_dim,
                 "set_model_dims: inconsistent dimensions");
    h->dims = *d;
    h->dims_set = 1;
    return IVLM_OK;
}

extern "C" size_t ivlm_sam_encode_arena_bytes(ivlm_handle h, int32_t N) {
    if (!h || !h->dims_set || N <= 0) return 0;
    const ivlm_model_dims& d = h->dims;
    const size_t g = d.sam_img / d.sam_patch, S = g * g, E = d.sam_embed_dim, nw = (g + d.sam_window - 1) / d.sam_window;
    const size_t rows = (size_t)N * S, wrows = (size_t)N * nw * nw * d.sam_window * d.sam_window;
    const size_t qkv_rows = wrows > rows ? wrows : rows;
    // x (two buffers), y, o: rows x E; qkv: qkv_rows x 3E; h: rows x max(4E, patch operand, 3x3 operand); slack for alignment
    size_t hw = 4 * E;
    const size_t kk = 3 * (size_t)d.sam_patch * d.sam_patch, k9 = 9 * (size_t)d.sam_out_chans;
    hw = hw > kk ? hw : kk;
    hw = hw > k9 ? hw : k9;
    const size_t Ew = E > (size_t)d.sam_out_chans ? E : (size_t)d.sam_out_chans;
    return 2 * (2 * rows * E + 2 * rows * Ew + qkv_rows * 3 * E + rows * hw) + 16 * 256;
}

// images [N,3,S,S] bf16 -> emb [N, g*g, out_chans] bf16 token-major (image_encoder.py:110-125).  Window blocks run on the real
// tokens only (win_inv / win_pads / win_map as model.py builds them); weights "sam.*" as bound by model.py's _Weights.
extern "C" int ivlm_sam_encode(ivlm_handle h, const ivlm_sam_encode_args* a, void* stream) {
    IVLM_REQUIRE(h && a && a->images && a->emb && a->arena && a->N > 0, "sam_encode: bad arguments");
    IVLM_REQUIRE(h->dims_set, "sam_encode: call ivlm_set_model_dims first");
    const ivlm_model_dims& d = h->dims;
    const int g = d.sam_img / d.sam_patch, S = g * g, E = d.sam_embed_dim, nh = d.sam_heads, hd = E / nh, ws = d.sam_window;
    const int nw = (g + ws - 1) / ws, N = a->N, O = d.sam_out_chans;
    const int rows = N * S, Bw = N * nw * nw, wrows = Bw * ws * ws;
    IVLM_REQUIRE(hd == 80 && ws == 14 && a->win_map && a->win_inv, "sam_encode: the stage driver covers the ViT-H geometry (head_dim 80, 14x14 windows)");
    Arena arena(a->arena, a->arena_bytes);
    const int kk = 3 * d.sam_patch * d.sam_patch;
    size_t hw = 4 * (size_t)E;
    hw = hw > (size_t)kk ? hw : (size_t)kk;
    hw = hw > (size_t)9 * O ? hw : (size_t)9 * O;
    const size_t Ew = E > O ? E : O;
    IVLM_TAKE(xa, uint16_t, (size_t)rows * E);
    IVLM_TAKE(xb, uint16_t, (size_t)rows * E);
    IVLM_TAKE(y, uint16_t, (size_t)rows * Ew);
    IVLM_TAKE(o, uint16_t, (size_t)rows * Ew);
    IVLM_TAKE(qkv, uint16_t, (size_t)(wrows > rows ? wrows : rows) * 3 * E);
    IVLM_TAKE(hbuf, uint16_t, (size_t)rows * hw);
    IVLM_W(w_patch, "sam.w_patch"); IVLM_W(b_patch, "sam.b_patch"); IVLM_W(pos, "sam.pos");
    IVLM_REQUIRE(kk % 8 == 0, "sam_encode: patch operand width %d is not a multiple of 8", kk);
    IVLM_TRY(ivlm_im2col_patch_bf16(h, a->images, hbuf, N, 3, d.sam_img, d.sam_img, d.sam_patch, kk, stream));
    IVLM_TRY(gemm(h, hbuf, kk, w_patch, kk, xa, E, rows, E, kk, b_patch, 0, pos, E, nullptr, S, -1, IVLM_BF16, stream));
    uint16_t *x = xa, *xn = xb;
    for (int i = 0; i < d.sam_depth; ++i) {
        const std::string p = "sam.blocks." + std::to_string(i) + ".";
        IVLM_W(n1g, p + "n1g"); IVLM_W(n1b, p + "n1b"); IVLM_W(wqkv, p + "wqkv"); IVLM_W(bqkv, p + "bqkv");
        IVLM_W(rph, p + "rph"); IVLM_W(rpw, p + "rpw"); IVLM_W(wo, p + "wo"); IVLM_W(bo, p + "bo");
        IVLM_W(n2g, p + "n2g"); IVLM_W(n2b, p + "n2b"); IVLM_W(w1, p + "w1"); IVLM_W(b1, p + "b1");
        IVLM_W(w2, p + "w2"); IVLM_W(b2, p + "b2");
        const bool global = (d.sam_global_mask >> i) & 1;
        IVLM_TRY(ivlm_layernorm_bf16(h, x, y, n1g, n1b, rows, E, 1e-6f, nullptr, 0, stream));
        if (global) {
            IVLM_TRY(gemm(h, y, E, wqkv, E, qkv, 3 * E, rows, 3 * E, E, bqkv, 0, nullptr, 0, nullptr, 0, -1, IVLM_BF16, stream));
            IVLM_TRY(ivlm_sam_attention_bf16(h, qkv, rph, rpw, o, N, nh, g, g, hd, E, nullptr, stream));
        } else {
            if (a->n_pads > 0) IVLM_TRY(ivlm_fill_rows_bf16(h, qkv, 3 * E, a->win_pads, a->n_pads, bqkv, 3 * E, stream));
            IVLM_TRY(gemm(h, y, E, wqkv, E, qkv, 3 * E, rows, 3 * E, E, bqkv, 0, nullptr, 0, a->win_inv, 0, -1, IVLM_BF16, stream));
            IVLM_TRY(ivlm_sam_attention_bf16(h, qkv, rph, rpw, o, Bw, nh, ws, ws, hd, E, a->win_map, stream));
        }
        IVLM_TRY(gemm(h, o, E, wo, E, xn, E, rows, E, E, bo, 0, x, E, nullptr, 0, -1, IVLM_BF16, stream));
        IVLM_TRY(ivlm_layernorm_bf16(h, xn, y, n2g, n2b, rows, E, 1e-6f, nullptr, 0, stream));
        IVLM_TRY(gemm(h, y, E, w1, E, hbuf, 4 * E, rows, 4 * E, E, b1, IVLM_ACT_GELU, nullptr, 0, nullptr, 0, -1, IVLM_BF16, stream));
        IVLM_TRY(gemm(h, hbuf, 4 * E, w2, 4 * E, x, E, rows, E, 4 * E, b2, 0, xn, E, nullptr, 0, -1, IVLM_BF16, stream));
    }
    IVLM_W(neck0, "sam.neck0"); IVLM_W(n1g, "sam.n1g"); IVLM_W(n1b, "sam.n1b"); IVLM_W(neck2, "sam.neck2");
    IVLM_W(n3g, "sam.n3g"); IVLM_W(n3b, "sam.n3b");
    uint16_t *t0 = y, *t1 = o;   // [rows, O] temporaries inside the dead E-wide buffers
    IVLM_TRY(gemm(h, x, E, neck0, E, t0, O, rows, O, E, nullptr, 0, nullptr, 0, nullptr, 0, -1, IVLM_BF16, stream));
    IVLM_TRY(ivlm_layernorm_bf16(h, t0, t1, n1g, n1b, rows, O, 1e-6f, nullptr, 0, stream));
    IVLM_TRY(ivlm_im2col_3x3_bf16(h, t1, hbuf, N, g, g, O, stream));
    IVLM_TRY(gemm(h, hbuf, 9 * O, neck2, 9 * O, t0, O, rows, O, 9 * O, nullptr, 0, nullptr, 0, nullptr, 0, -1, IVLM_BF16, stream));
    IVLM_TRY(ivlm_layernorm_bf16(h, t0, a->emb, n3g, n3b, rows, O, 1e-6f, nullptr, 0, stream));
    return IVLM_OK;
}

extern "C" size_t ivlm_llm_arena_bytes(ivlm_handle h, int32_t tokens) {
    if (!h || !h->dims_set || tokens <= 0) return 0;
    const ivlm_model_dims& d = h->dims;
    const size_t T = tokens, D = d.llm_hidden, F = d.llm_intermediate;
    // x (two), y, q, k, v, o: T x D; qkv: T x 3D; gate-up: T x 2F; act: T x F; logits: 64 x vocab fp32
    return 2 * (8 * T * D + 3 * T * D + 2 * T * F + T * F) + 4 * (size_t)64 * (d.llm_vocab + 8) + 32 * 256;   // 8th T x D: the last-row gather
}

// embeds [B*S, D] bf16 (B sequences of S rows, right-padded) -> K/V pages, hidden [B, max_len, D] (rows [0,S) of every
// sequence = normed last-layer states), next_tok [B] = greedy token after each sequence's last valid row (last_rows [B] flat
// row indices, or NULL for row S-1).  HF LlamaModel + lm_head (llava_llama.py:93-105), weights "llm.*".
extern "C" int ivlm_llm_prefill(ivlm_handle h, const ivlm_llm_prefill_args* a, void* stream) {
    IVLM_REQUIRE(h && a && a->embeds && a->positions && a->slot_map && a->k_cache && a->v_cache && a->hidden && a->next_tok && a->arena &&
                     a->B > 0 && a->S > 0, "llm_prefill: bad arguments");
    IVLM_REQUIRE(h->dims_set, "llm_prefill: call ivlm_set_model_dims first");
    const ivlm_model_dims& d = h->dims;
    const int D = d.llm_hidden, F = d.llm_intermediate, nh = d.llm_heads, hd = d.llm_head_dim, T = a->B * a->S;
    IVLM_CHECK_CUDA(cudaGetLastError());   // a launch error left behind by earlier work on this device would otherwise surface below
    Arena arena(a->arena, a->arena_bytes);
    IVLM_TAKE(xa, uint16_t, (size_t)T * D);
    IVLM_TAKE(xb, uint16_t, (size_t)T * D);
    IVLM_TAKE(y, uint16_t, (size_t)T * D);
    IVLM_TAKE(qkv, uint16_t, (size_t)T * 3 * D);
    IVLM_TAKE(q, uint16_t, (size_t)T * D);
    IVLM_TAKE(k, uint16_t, (size_t)T * D);
    IVLM_TAKE(v, uint16_t, (size_t)T * D);
    IVLM_TAKE(o, uint16_t, (size_t)T * D);
    IVLM_TAKE(gu, uint16_t, (size_t)T * 2 * F);
    IVLM_TAKE(act, uint16_t, (size_t)T * F);
    IVLM_W(cos_t, "llm.rope_cos"); IVLM_W(sin_t, "llm.rope_sin");
    // layer input x (the embeddings, then xb); x1 = x + attention -> xa; x2 = x1 + mlp -> xb (its old content, the layer input,
    // is dead once x1 exists -- stream order)
    const uint16_t* x = reinterpret_cast<const uint16_t*>(a->embeds);
    for (int i = 0; i < d.llm_layers; ++i) {
        const std::string p = "llm." + std::to_string(i) + ".";
        IVLM_W(ln1, p + "ln1"); IVLM_W(ln2, p + "ln2"); IVLM_W(wqkv, p + "wqkv"); IVLM_W(wo, p + "wo"); IVLM_W(wgu, p + "wgu");
        IVLM_W(wd, p + "wd");
        IVLM_TRY(ivlm_rmsnorm_bf16(h, x, y, ln1, T, D, d.llm_rms_eps, stream));
        IVLM_TRY(gemm(h, y, D, wqkv, D, qkv, 3 * D, T, 3 * D, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
        IVLM_TRY(ivlm_rope_kv_store_bf16(h, qkv, a->positions, a->slot_map, cos_t, sin_t, q, k, v, a->k_cache[i], a->v_cache[i], T, nh, hd,
                                         a->page_size, d.llm_paired_layout, stream));
        ivlm_attn_args at;
        memset(&at, 0, sizeof(at));
        at.q = q; at.k = k; at.v = v; at.out = o;
        at.q_bs = at.k_bs = at.v_bs = at.o_bs = (int64_t)a->S * D;
        at.q_ts = at.k_ts = at.v_ts = at.o_ts = D;
        at.q_hs = at.k_hs = at.v_hs = at.o_hs = hd;
        at.B = a->B; at.H = nh; at.Sq = a->S; at.Sk = a->S; at.D = hd;
        at.scale = 1.0f / sqrtf((float)hd);
        at.causal = 1;
        IVLM_TRY(ivlm_attention_bf16(h, &at, stream));
        IVLM_TRY(gemm(h, o, D, wo, D, xa, D, T, D, D, nullptr, 0, x, D, nullptr, 0, 0, IVLM_BF16, stream));
        IVLM_TRY(ivlm_rmsnorm_bf16(h, xa, y, ln2, T, D, d.llm_rms_eps, stream));
        if (d.llm_paired_layout && T > 64 && 2 * F > 32 && F % 8 == 0) {
            // gate / up rows are interleaved: the GEMM epilogue applies the SwiGLU gate and writes [T, F] directly
            IVLM_TRY(gemm(h, y, D, wgu, D, act, F, T, 2 * F, D, nullptr, IVLM_ACT_SWIGLU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
        } else {
            IVLM_TRY(gemm(h, y, D, wgu, D, gu, 2 * F, T, 2 * F, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
            IVLM_TRY(ivlm_silu_mul_bf16(h, gu, act, T, F, d.llm_paired_layout, stream));
        }
        IVLM_TRY(gemm(h, act, F, wd, F, xb, D, T, D, F, nullptr, 0, xa, D, nullptr, 0, 0, IVLM_BF16, stream));
        x = xb;
    }
    IVLM_W(norm, "llm.norm");
    // normed states of every row -> hidden[b, 0:S, :]
    IVLM_TRY(ivlm_rmsnorm_bf16(h, x, y, norm, T, D, d.llm_rms_eps, stream));
    IVLM_REQUIRE(a->max_len >= a->S, "llm_prefill: hidden holds %d rows per sequence, the prompt has %d", a->max_len, a->S);
    IVLM_CHECK_CUDA(cudaMemcpy2DAsync(a->hidden, sizeof(uint16_t) * (size_t)a->max_len * D, y, sizeof(uint16_t) * (size_t)a->S * D,
                                      sizeof(uint16_t) * (size_t)a->S * D, (size_t)a->B, cudaMemcpyDeviceToDevice,
                                      reinterpret_cast<cudaStream_t>(stream)));
    // greedy token after each sequence's last valid row
    IVLM_W(lm_head, "llm.lm_head");
    IVLM_TAKE(last, uint16_t, (size_t)a->B * D);
    if (a->last_rows != nullptr) {
        IVLM_TRY(ivlm_gather_rows_bf16(h, y, a->last_rows, last, a->B, D, stream));
    } else {
        for (int b = 0; b < a->B; ++b)
            IVLM_CHECK_CUDA(cudaMemcpyAsync(last + (size_t)b * D, y + ((size_t)b * a->S + a->S - 1) * D, sizeof(uint16_t) * D,
                                            cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
    }
    const int vocab_ld = d.llm_vocab;
    for (int b0 = 0; b0 < a->B; b0 += 64) {
        const int nb = a->B - b0 < 64 ? a->B - b0 : 64;
        IVLM_TAKE(logits, float, (size_t)nb * vocab_ld);
        IVLM_TRY(gemm(h, last + (size_t)b0 * D, D, lm_head, D, logits, vocab_ld, nb, d.llm_vocab, D, nullptr, 0, nullptr, 0, nullptr, 0, 1,
                      IVLM_F32, stream));
        IVLM_TRY(ivlm_argmax_f32(h, logits, a->next_tok + b0, nb, d.llm_vocab, vocab_ld, stream));
    }
    return IVLM_OK;
}

// One token per sequence through the paged KV cache: the fixed launch sequence a CUDA graph captures (model.py llm_decode_step):
// device-side bookkeeping, 5 launches per layer (ivlm_decode_linear) for B <= 8, final norm, hidden-state filing, lm_head, argmax.
extern "C" int ivlm_llm_decode_step(ivlm_handle h, const ivlm_llm_decode_args* a, void* stream) {
    IVLM_REQUIRE(h && a && a->state && a->next && a->done && a->out_tokens && a->tok && a->pos && a->slot && a->seq_lens && a->slot_base &&
                     a->k_cache && a->v_cache && a->block_table && a->hidden && a->hid_step && a->arena && a->B > 0 && a->B <= 8,
                 "llm_decode_step: bad arguments (the stage driver covers B <= 8; larger batches use the op-level path)");
    IVLM_REQUIRE(h->dims_set && h->dims.llm_paired_layout, "llm_decode_step: needs ivlm_set_model_dims with the paired / interleaved layout");
    const ivlm_model_dims& d = h->dims;
    const int D = d.llm_hidden, F = d.llm_intermediate, nh = d.llm_heads, hd = d.llm_head_dim, B = a->B;
    Arena arena(a->arena, a->arena_bytes);
    IVLM_TAKE(xa, uint16_t, (size_t)B * D);
    IVLM_TAKE(xb, uint16_t, (size_t)B * D);
    IVLM_TAKE(q, uint16_t, (size_t)B * D);
    IVLM_TAKE(o, uint16_t, (size_t)B * D);
    IVLM_TAKE(act, uint16_t, (size_t)B * F);
    IVLM_TAKE(logits, float, (size_t)B * d.llm_vocab);
    IVLM_W(embed, "llm.embed"); IVLM_W(cos_t, "llm.rope_cos"); IVLM_W(sin_t, "llm.rope_sin");
    IVLM_TRY(ivlm_decode_prepare(h, a->state, a->S, a->S_rows, a->scripted, a->G, a->next, a->done, a->out_tokens, a->tok, a->pos, a->slot,
                                 a->seq_lens, a->slot_base, a->eos, a->pad, B, stream));
    IVLM_TRY(ivlm_embed_gather_bf16(h, embed, a->tok, xa, B, D, d.llm_vocab, stream));
    uint16_t *x = xa, *xn = xb;
    ivlm_decode_linear_args g;
    for (int i = 0; i < d.llm_layers; ++i) {
        const std::string p = "llm." + std::to_string(i) + ".";
        IVLM_W(ln1, p + "ln1"); IVLM_W(ln2, p + "ln2"); IVLM_W(wqkv, p + "wqkv"); IVLM_W(wo, p + "wo"); IVLM_W(wgu, p + "wgu");
        IVLM_W(wd, p + "wd");
        memset(&g, 0, sizeof(g));
        g.a = x; g.lda = D; g.w = wqkv; g.ldw = D; g.M = B; g.N = 3 * D; g.K = D; g.norm_gamma = ln1; g.norm_eps = d.llm_rms_eps;
        g.epilogue = IVLM_EPI_ROPE_KV; g.out = q; g.ldo = D; g.out_dtype = IVLM_BF16;
        g.positions = a->pos; g.slot_map = a->slot; g.cos_t = cos_t; g.sin_t = sin_t; g.k_cache = a->k_cache[i]; g.v_cache = a->v_cache[i];
        g.H = nh; g.hd = hd; g.page_size = a->page_size;
        g.prefetch_w = wo; g.prefetch_ldw = D; g.prefetch_N = D; g.prefetch_K = D; g.prefetch_stages = 64;   // all of o_proj: it loads under the attention launch
        IVLM_TRY(ivlm_decode_linear(h, &g, stream));
        IVLM_TRY(ivlm_decode_attention_paged_bf16(h, q, a->k_cache[i], a->v_cache[i], a->block_table, a->seq_lens, o, B, nh, hd, a->page_size,
                                                  a->max_pages, 1.0f / sqrtf((float)hd), stream));
        memset(&g, 0, sizeof(g));
        g.a = o; g.lda = D; g.w = wo; g.ldw = D; g.M = B; g.N = D; g.K = D; g.epilogue = IVLM_EPI_PLAIN; g.residual = x; g.ldr = D;
        g.out = xn; g.ldo = D; g.out_dtype = IVLM_BF16;
        g.prefetch_w = wgu; g.prefetch_ldw = D; g.prefetch_N = 2 * F; g.prefetch_K = D;
        IVLM_TRY(ivlm_decode_linear(h, &g, stream));
        memset(&g, 0, sizeof(g));
        g.a = xn; g.lda = D; g.w = wgu; g.ldw = D; g.M = B; g.N = 2 * F; g.K = D; g.norm_gamma = ln2; g.norm_eps = d.llm_rms_eps;
        g.epilogue = IVLM_EPI_SWIGLU; g.out = act; g.ldo = F; g.out_dtype = IVLM_BF16;
        g.prefetch_w = wd; g.prefetch_ldw = F; g.prefetch_N = D; g.prefetch_K = F;
        IVLM_TRY(ivlm_decode_linear(h, &g, stream));
        memset(&g, 0, sizeof(g));
        g.a = act; g.lda = F; g.w = wd; g.ldw = F; g.M = B; g.N = D; g.K = F; g.epilogue = IVLM_EPI_PLAIN; g.residual = xn; g.ldr = D;
        g.out = x; g.ldo = D; g.out_dtype = IVLM_BF16;
        if (i + 1 < d.llm_layers) {
            IVLM_W(wqkv_next, "llm." + std::to_string(i + 1) + ".wqkv");
            g.prefetch_w = wqkv_next; g.prefetch_ldw = D; g.prefetch_N = 3 * D; g.prefetch_K = D;
        }
        IVLM_TRY(ivlm_decode_linear(h, &g, stream));
    }
    IVLM_W(norm, "llm.norm"); IVLM_W(lm_head, "llm.lm_head");
    IVLM_TRY(ivlm_rmsnorm_bf16(h, x, a->hid_step, norm, B, D, d.llm_rms_eps, stream));
    IVLM_TRY(ivlm_decode_finish(h, a->state, a->S, a->S_rows, a->hid_step, a->hidden, B, D, a->max_len, stream));
    memset(&g, 0, sizeof(g));
    g.a = a->hid_step; g.lda = D; g.w = lm_head; g.ldw = D; g.M = B; g.N = d.llm_vocab; g.K = D; g.epilogue = IVLM_EPI_PLAIN;
    g.out = logits; g.ldo = d.llm_vocab; g.out_dtype = IVLM_F32;
    IVLM_TRY(ivlm_decode_linear(h, &g, stream));
    IVLM_TRY(ivlm_argmax_f32(h, logits, a->next, B, d.llm_vocab, d.llm_vocab, stream));
    return IVLM_OK;
}

// ------------------------------------------------------------------------------------------------ CLIP tower + projector
extern "C" size_t ivlm_clip_encode_arena_bytes(ivlm_handle h, int32_t B) {
    if (!h || !h->dims_set || B <= 0 || h->dims.clip_hidden <= 0) return 0;
    const ivlm_model_dims& d = h->dims;
    const size_t g = d.clip_img / d.clip_patch, T = g * g + 1, C = d.clip_hidden, rows = (size_t)B * T;
    const Weight* w1 = find_weight(h, "clip.0.w1");
    const size_t F = w1 ? (size_t)w1->shape[0] : 4 * C;   // mlp width
    // cols [B*(T-1), ldk]; h (two), y, o: rows x C; qkv: rows x 3C; mlp: rows x F; patches: B*(T-1) x C
    return 2 * ((size_t)B * (T - 1) * d.clip_ldk + 4 * rows * C + 3 * rows * C + F * rows + (size_t)B * (T - 1) * C) + 16 * 256;
}

// CLIPVisionTower.forward + feature_select('patch', layer -2) + mm_projector (clip_encoder.py:31-60, llava_arch.py:93-96):
// images [B,3,S,S] bf16 -> feats [B, T-1, llm_hidden] bf16.  The launch sequence of model.py's _Engine.clip_encode; weights
// "clip.*" / "mm.*"; patch_rows [B*(T-1)] = b*T + 1 + j (destination row of every patch token), cls_rows [B] = b*T.
extern "C" int ivlm_clip_encode(ivlm_handle h, const ivlm_clip_encode_args* a, void* stream) {
    IVLM_REQUIRE(h && a && a->images && a->feats && a->patch_rows && a->cls_rows && a->arena && a->B > 0, "clip_encode: bad arguments");
    IVLM_REQUIRE(h->dims_set && h->dims.clip_hidden > 0, "clip_encode: call ivlm_set_model_dims (with the clip_* fields) first");
    const ivlm_model_dims& d = h->dims;
    const int B = a->B, g = d.clip_img / d.clip_patch, T = g * g + 1, C = d.clip_hidden, nh = d.clip_heads, hd = C / nh, ldk = d.clip_ldk;
    const int rows = B * T, prow = B * (T - 1);
    Arena arena(a->arena, a->arena_bytes);
    IVLM_TAKE(cols, uint16_t, (size_t)prow * ldk);
    IVLM_TAKE(ha, uint16_t, (size_t)rows * C);
    IVLM_TAKE(hb, uint16_t, (size_t)rows * C);
    IVLM_TAKE(y, uint16_t, (size_t)rows * C);
    IVLM_TAKE(o, uint16_t, (size_t)rows * C);
    IVLM_TAKE(qkv, uint16_t, (size_t)rows * 3 * C);
    const Weight* w1_first = find_weight(h, "clip.0.w1");
    IVLM_REQUIRE(w1_first != nullptr, "clip_encode: weights 'clip.*' are not bound (ivlm_bind_weights)");
    const int F = (int)w1_first->shape[0];   // mlp width (4 C for the OpenAI towers)
    IVLM_TAKE(mlp, uint16_t, (size_t)rows * F);
    IVLM_TAKE(patches, uint16_t, (size_t)prow * C);
    IVLM_W(w_patch, "clip.w_patch"); IVLM_W(pos, "clip.pos"); IVLM_W(cls_pos, "clip.cls_pos"); IVLM_W(pre_g, "clip.pre_g"); IVLM_W(pre_b, "clip.pre_b");
    IVLM_TRY(ivlm_im2col_patch_bf16(h, a->images, cols, B, 3, d.clip_img, d.clip_img, d.clip_patch, ldk, stream));
    IVLM_TRY(ivlm_fill_rows_bf16(h, ha, C, a->cls_rows, B, cls_pos, C, stream));
    IVLM_TRY(gemm(h, cols, ldk, w_patch, ldk, ha, C, prow, C, ldk, nullptr, 0, pos, C, a->patch_rows, T, -1, IVLM_BF16, stream));
    IVLM_TRY(ivlm_layernorm_bf16(h, ha, hb, pre_g, pre_b, rows, C, d.clip_eps, nullptr, 0, stream));
    uint16_t *x = hb, *xn = ha;
    for (int i = 0; i < d.clip_layers; ++i) {
        const std::string p = "clip." + std::to_string(i) + ".";
        IVLM_W(ln1g, p + "ln1g"); IVLM_W(ln1b, p + "ln1b"); IVLM_W(wqkv, p + "wqkv"); IVLM_W(bqkv, p + "bqkv"); IVLM_W(wo, p + "wo");
        IVLM_W(bo, p + "bo"); IVLM_W(ln2g, p + "ln2g"); IVLM_W(ln2b, p + "ln2b"); IVLM_W(w1, p + "w1"); IVLM_W(b1, p + "b1");
        IVLM_W(w2, p + "w2"); IVLM_W(b2, p + "b2");
        IVLM_TRY(ivlm_layernorm_bf16(h, x, y, ln1g, ln1b, rows, C, d.clip_eps, nullptr, 0, stream));
        IVLM_TRY(gemm(h, y, C, wqkv, C, qkv, 3 * C, rows, 3 * C, C, bqkv, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
        ivlm_attn_args at;
        memset(&at, 0, sizeof(at));
        at.q = qkv; at.k = qkv + C; at.v = qkv + 2 * C; at.out = o;
        at.q_bs = at.k_bs = at.v_bs = (int64_t)T * 3 * C; at.q_ts = at.k_ts = at.v_ts = 3 * C; at.q_hs = at.k_hs = at.v_hs = hd;
        at.o_bs = (int64_t)T * C; at.o_ts = C; at.o_hs = hd;
        at.B = B; at.H = nh; at.Sq = T; at.Sk = T; at.D = hd;
        at.scale = 1.0f / sqrtf((float)hd);
        IVLM_TRY(ivlm_attention_bf16(h, &at, stream));
        IVLM_TRY(gemm(h, o, C, wo, C, xn, C, rows, C, C, bo, 0, x, C, nullptr, 0, 0, IVLM_BF16, stream));
        IVLM_TRY(ivlm_layernorm_bf16(h, xn, y, ln2g, ln2b, rows, C, d.clip_eps, nullptr, 0, stream));
        IVLM_TRY(gemm(h, y, C, w1, C, mlp, F, rows, F, C, b1, IVLM_ACT_QUICK_GELU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
        IVLM_TRY(gemm(h, mlp, F, w2, F, x, C, rows, C, F, b2, 0, xn, C, nullptr, 0, 0, IVLM_BF16, stream));
    }
    IVLM_W(mm_w, "mm.w"); IVLM_W(mm_b, "mm.b");
    IVLM_TRY(ivlm_gather_rows_bf16(h, x, a->patch_rows, patches, prow, C, stream));
    IVLM_TRY(gemm(h, patches, C, mm_w, C, a->feats, d.llm_hidden, prow, d.llm_hidden, C, mm_b, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    return IVLM_OK;
}

// ------------------------------------------------------------------------------------------------ [SEG] head
// text_hidden_fcs[0] on the rows that predict [SEG] (InteractVLM.py:100-112,551-556) and the camera conditioning of
// process_embeddings (InteractVLM.py:268-294) for the released configurations: VIv1CamPoseEncoder (components.py:541-572) or no
// conditioning.  hidden_rows [n, llm_hidden], cam [n,V,5] -> prompt [n,V,out] (+ emb [n,out], the un-gated embedding).
extern "C" int ivlm_seg_head(ivlm_handle h, const void* hidden_rows, const void* cam, void* prompt, void* emb, int32_t n, int32_t V,
                             void* arena_, size_t arena_bytes, void* stream) {
    IVLM_REQUIRE(h && hidden_rows && prompt && emb && arena_ && n > 0 && V > 0, "seg_head: bad arguments");
    IVLM_REQUIRE(h->dims_set, "seg_head: call ivlm_set_model_dims first");
    const int D = h->dims.llm_hidden;
    IVLM_W(fc0_w, "seg.fc0_w"); IVLM_W(fc0_b, "seg.fc0_b"); IVLM_W(fc2_w, "seg.fc2_w"); IVLM_W(fc2_b, "seg.fc2_b");
    const int Hm = (int)fc0_w_w->shape[0], O = (int)fc2_w_w->shape[0];
    Arena arena(arena_, arena_bytes);
    IVLM_TAKE(y, uint16_t, (size_t)n * Hm);
    IVLM_TRY(gemm(h, hidden_rows, D, fc0_w, D, y, Hm, n, Hm, D, fc0_b, IVLM_ACT_RELU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    IVLM_TRY(gemm(h, y, Hm, fc2_w, Hm, emb, O, n, O, Hm, fc2_b, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    if (find_weight(h, "seg.cam.w1") != nullptr) {
        IVLM_REQUIRE(cam != nullptr, "seg_head: camera parameters missing");
        IVLM_W(w1, "seg.cam.w1"); IVLM_W(b1, "seg.cam.b1"); IVLM_W(w2, "seg.cam.w2"); IVLM_W(b2, "seg.cam.b2"); IVLM_W(wv, "seg.cam.wv");
        IVLM_W(bv, "seg.cam.bv");
        IVLM_REQUIRE(wv_w->shape[0] == V && O == 256, "seg_head: the camera gate covers %d views of 256 channels", (int)wv_w->shape[0]);
        IVLM_TRY(ivlm_cam_gate_bf16(h, cam, emb, w1, b1, w2, b2, wv, bv, prompt, n, V, stream));
    } else {
        // no conditioning: every view gets the embedding
        IVLM_CHECK_CUDA(cudaMemcpy2DAsync(prompt, sizeof(uint16_t) * (size_t)V * O, emb, sizeof(uint16_t) * (size_t)O, sizeof(uint16_t) * (size_t)O,
                                          (size_t)n, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
        for (int v = 1; v < V; ++v)
            IVLM_CHECK_CUDA(cudaMemcpy2DAsync(reinterpret_cast<uint16_t*>(prompt) + (size_t)v * O, sizeof(uint16_t) * (size_t)V * O, emb,
                                              sizeof(uint16_t) * (size_t)O, sizeof(uint16_t) * (size_t)O, (size_t)n, cudaMemcpyDeviceToDevice,
                                              reinterpret_cast<cudaStream_t>(stream)));
    }
    return IVLM_OK;
}

// ------------------------------------------------------------------------------------------------ prompt encoder + mask decoder
namespace ivlm {
struct DecAttnW { const Weight *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo; };
static bool dec_attn_weights(ivlm_ctx* h, const std::string& p, DecAttnW& w) {
    w.wq = find_weight(h, p + "wq"); w.bq = find_weight(h, p + "bq"); w.wk = find_weight(h, p + "wk"); w.bk = find_weight(h, p + "bk");
    w.wv = find_weight(h, p + "wv"); w.bv = find_weight(h, p + "bv"); w.wo = find_weight(h, p + "wo"); w.bo = find_weight(h, p + "bo");
    return w.wq && w.bq && w.wk && w.bk && w.wv && w.bv && w.wo && w.bo;
}
// transformer.py:185-242 Attention: q [B,Nq,C], k / v [B,Nk,C] -> out [B,Nq,C] (+ residual); qp / kp / vp / ao: scratch
static int dec_attn(ivlm_ctx* h, const DecAttnW& w, const void* q, const void* k, const void* v, const void* residual, void* out, int B, int Nq,
                    int Nk, int C, int heads, uint16_t* qp, uint16_t* kp, uint16_t* vp, uint16_t* ao, void* stream) {
    const int I = (int)w.wq->shape[0];   // internal width (C, or C / downsample_rate for the cross attentions)
    IVLM_TRY(gemm(h, q, C, w.wq->ptr, C, qp, I, B * Nq, I, C, w.bq->ptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    IVLM_TRY(gemm(h, k, C, w.wk->ptr, C, kp, I, B * Nk, I, C, w.bk->ptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    IVLM_TRY(gemm(h, v, C, w.wv->ptr, C, vp, I, B * Nk, I, C, w.bv->ptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
    IVLM_TRY(ivlm_attn_small_bf16(h, qp, kp, vp, ao, B, 0, Nq, Nk, heads, I / heads, stream));
    return gemm(h, ao, I, w.wo->ptr, I, out, C, B * Nq, C, I, w.bo->ptr, 0, residual, C, nullptr, 0, 0, IVLM_BF16, stream);
}
}  // namespace ivlm

extern "C" size_t ivlm_mask_decode_arena_bytes(ivlm_handle h, int32_t n, int32_t V) {
    if (!h || !h->dims_set || n <= 0 || V <= 0) return 0;
    const ivlm_model_dims& d = h->dims;
    const size_t g = d.sam_img / d.sam_patch, S = g * g, C = d.sam_out_chans, nv = (size_t)n * V, ntok = 5 + (size_t)V;
    // image side: keys (two), k, kp, vp, ao(image->token out), up1 (two): 8 buffers of nv*S*C; token side: a handful of nv*ntok*2048
    return 2 * (8 * nv * S * C + 12 * nv * ntok * 2048) + (size_t)(5 + nv) * C * 2 + 64 * 256;
}

// ModifiedSAM.forward -> PromptEncoder (text tokens as sparse prompts, no_mask dense embedding, dense PE) -> MaskDecoder.predict_masks
// (InteractVLM.py:40-63, prompt_encoder.py:140-186, mask_decoder.py:116-164, transformer.py:62-182): emb [n*V, S, C] token-major,
// prompt [n, V, C] -> low-res logits [n*V, 4g, 4g] fp32.  Every view of a sample sees the 5 output tokens + the sample's V prompt
// tokens.  The launch sequence of model.py's _Engine.mask_decode; weights "dec.*"; tok_idx [n*V*(5+V)]: row of the token table
// [out_tokens (5 rows); prompt (n*V rows)] that makes up token row r.
extern "C" int ivlm_mask_decode(ivlm_handle h, const ivlm_mask_decode_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && a && a->emb && a->prompt && a->lowres && a->tok_idx && a->arena && a->n > 0 && a->V > 0 && a->heads > 0,
                 "mask_decode: bad arguments");
    IVLM_REQUIRE(h->dims_set, "mask_decode: call ivlm_set_model_dims first");
    const ivlm_model_dims& d = h->dims;
    const int g = d.sam_img / d.sam_patch, S = g * g, C = d.sam_out_chans, n = a->n, V = a->V, nv = n * V, heads = a->heads;
    IVLM_W(out_tokens, "dec.out_tokens"); IVLM_W(no_mask, "dec.no_mask"); IVLM_W(dense_pe, "dec.dense_pe");
    const int n_out = (int)out_tokens_w->shape[0], ntok = n_out + V, QT = nv * ntok;
    IVLM_REQUIRE(n_out == 5 || n_out >= 1, "mask_decode: malformed output tokens");
    Arena arena(a->arena, a->arena_bytes);
    const size_t img = (size_t)nv * S * C, tk = (size_t)QT * 2048;
    IVLM_TAKE(table, uint16_t, (size_t)(n_out + nv) * C);
    IVLM_TAKE(tokens, uint16_t, (size_t)QT * C);
    IVLM_TAKE(qa, uint16_t, (size_t)QT * C);
    IVLM_TAKE(qb, uint16_t, (size_t)QT * C);
    IVLM_TAKE(qc, uint16_t, (size_t)QT * C);
    IVLM_TAKE(qpe, uint16_t, (size_t)QT * C);
    IVLM_TAKE(tq, uint16_t, tk);      // token-side projections / mlp hidden
    IVLM_TAKE(tk_, uint16_t, tk);
    IVLM_TAKE(tv, uint16_t, tk);
    IVLM_TAKE(ta, uint16_t, tk);
    IVLM_TAKE(keys_a, uint16_t, img);
    IVLM_TAKE(keys_b, uint16_t, img);
    IVLM_TAKE(kpe, uint16_t, img);
    IVLM_TAKE(ip, uint16_t, img);     // image-side projections (internal width <= C)
    IVLM_TAKE(iv, uint16_t, img);
    IVLM_TAKE(ia, uint16_t, img);
    IVLM_TAKE(up_a, uint16_t, img);
    IVLM_TAKE(up_b, uint16_t, img);
    // token table -> tokens [nv, ntok, C] (the query positional encoding of the decoder)
    IVLM_CHECK_CUDA(cudaMemcpyAsync(table, out_tokens, sizeof(uint16_t) * (size_t)n_out * C, cudaMemcpyDeviceToDevice, stream));
    IVLM_CHECK_CUDA(cudaMemcpyAsync(table + (size_t)n_out * C, a->prompt, sizeof(uint16_t) * (size_t)nv * C, cudaMemcpyDeviceToDevice, stream));
    IVLM_TRY(ivlm_gather_rows_bf16(h, table, a->tok_idx, tokens, QT, C, stream_));
    IVLM_TRY(ivlm_add_bcast_bf16(h, a->emb, no_mask, keys_a, (int64_t)img, C, stream_));
    uint16_t *keys = keys_a, *keys_n = keys_b;
    const uint16_t* queries = tokens;
    uint16_t* qbuf[3] = {qa, qb, qc};
    int qi = 0;
    auto next_q = [&]() { uint16_t* r = qbuf[qi]; qi = (qi + 1) % 3; return r; };
    int depth = 0;
    while (find_weight(h, "dec." + std::to_string(depth) + ".w1") != nullptr) ++depth;
    IVLM_REQUIRE(depth > 0, "mask_decode: no decoder layers bound (dec.0.*)");
    for (int i = 0; i < depth; ++i) {
        const std::string p = "dec." + std::to_string(i) + ".";
        DecAttnW sa, t2i, i2t;
        IVLM_REQUIRE(dec_attn_weights(h, p + "self.", sa) && dec_attn_weights(h, p + "t2i.", t2i) && dec_attn_weights(h, p + "i2t.", i2t),
                     "mask_decode: attention weights of layer %d are not bound", i);
        IVLM_W(n0g, p + "n0g"); IVLM_W(n0b, p + "n0b"); IVLM_W(n1g, p + "n1g"); IVLM_W(n1b, p + "n1b"); IVLM_W(n2g, p + "n2g");
        IVLM_W(n2b, p + "n2b"); IVLM_W(n3g, p + "n3g"); IVLM_W(n3b, p + "n3b"); IVLM_W(w1, p + "w1"); IVLM_W(b1, p + "b1");
        IVLM_W(w2, p + "w2"); IVLM_W(b2, p + "b2");
        const int Fm = (int)w1_w->shape[0];
        IVLM_REQUIRE(Fm <= 2048, "mask_decode: mlp width %d exceeds the scratch rows (2048)", Fm);
        uint16_t* t0 = next_q();
        if (i == 0) {
            IVLM_TRY(dec_attn(h, sa, queries, queries, queries, nullptr, t0, nv, ntok, ntok, C, heads, tq, tk_, tv, ta, stream_));
        } else {
            IVLM_TRY(ivlm_add_bcast_bf16(h, queries, tokens, qpe, (int64_t)QT * C, 0, stream_));
            IVLM_TRY(dec_attn(h, sa, qpe, qpe, queries, queries, t0, nv, ntok, ntok, C, heads, tq, tk_, tv, ta, stream_));
        }
        uint16_t* t1 = next_q();
        IVLM_TRY(ivlm_layernorm_bf16(h, t0, t1, n0g, n0b, QT, C, 1e-5f, nullptr, 0, stream_));
        IVLM_TRY(ivlm_add_bcast_bf16(h, t1, tokens, qpe, (int64_t)QT * C, 0, stream_));
        IVLM_TRY(ivlm_add_bcast_bf16(h, keys, dense_pe, kpe, (int64_t)img, (int64_t)S * C, stream_));
        uint16_t* t2 = next_q();   // == t0's buffer two steps later: t0 is dead
        IVLM_TRY(dec_attn(h, t2i, qpe, kpe, keys, t1, t2, nv, ntok, S, C, heads, tq, ip, iv, ta, stream_));
        uint16_t* t3 = next_q();
        IVLM_TRY(ivlm_layernorm_bf16(h, t2, t3, n1g, n1b, QT, C, 1e-5f, nullptr, 0, stream_));
        IVLM_TRY(gemm(h, t3, C, w1, C, tq, Fm, QT, Fm, C, b1, IVLM_ACT_RELU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream_));
        uint16_t* t4 = next_q();
        IVLM_TRY(gemm(h, tq, Fm, w2, Fm, t4, C, QT, C, Fm, b2, 0, t3, C, nullptr, 0, 0, IVLM_BF16, stream_));
        uint16_t* t5 = next_q();
        IVLM_TRY(ivlm_layernorm_bf16(h, t4, t5, n2g, n2b, QT, C, 1e-5f, nullptr, 0, stream_));
        IVLM_TRY(ivlm_add_bcast_bf16(h, t5, tokens, qpe, (int64_t)QT * C, 0, stream_));
        // image -> token: queries are the image tokens (+ pe), keys the decoder tokens (+ pe), values the decoder tokens
        IVLM_TRY(dec_attn(h, i2t, kpe, qpe, t5, keys, keys_n, nv, S, ntok, C, heads, ip, tk_, tv, ia, stream_));
        IVLM_TRY(ivlm_layernorm_bf16(h, keys_n, keys, n3g, n3b, (int64_t)nv * S, C, 1e-5f, nullptr, 0, stream_));
        queries = t5;
    }
    DecAttnW fin;
    IVLM_REQUIRE(dec_attn_weights(h, "dec.final.", fin), "mask_decode: final attention weights are not bound");
    IVLM_W(nfg, "dec.nfg"); IVLM_W(nfb, "dec.nfb");
    IVLM_TRY(ivlm_add_bcast_bf16(h, queries, tokens, qpe, (int64_t)QT * C, 0, stream_));
    IVLM_TRY(ivlm_add_bcast_bf16(h, keys, dense_pe, kpe, (int64_t)img, (int64_t)S * C, stream_));
    uint16_t* f0 = next_q();
    if (f0 == queries) f0 = next_q();
    IVLM_TRY(dec_attn(h, fin, qpe, kpe, keys, queries, f0, nv, ntok, S, C, heads, tq, ip, iv, ta, stream_));
    uint16_t* hs = next_q();
    if (hs == queries) hs = next_q();
    IVLM_TRY(ivlm_layernorm_bf16(h, f0, hs, nfg, nfb, QT, C, 1e-5f, nullptr, 0, stream_));
    // mask token 0 (row 1 of every view: row 0 is the IoU token) -> hypernetwork MLP
    uint16_t* mt = tq;
    IVLM_CHECK_CUDA(cudaMemcpy2DAsync(mt, sizeof(uint16_t) * (size_t)C, hs + C, sizeof(uint16_t) * (size_t)ntok * C, sizeof(uint16_t) * (size_t)C,
                                      (size_t)nv, cudaMemcpyDeviceToDevice, stream));
    IVLM_W(hw0, "dec.hyper0_w"); IVLM_W(hb0, "dec.hyper0_b"); IVLM_W(hw1, "dec.hyper1_w"); IVLM_W(hb1, "dec.hyper1_b");
    IVLM_W(hw2, "dec.hyper2_w"); IVLM_W(hb2, "dec.hyper2_b");
    const int Hh = (int)hw0_w->shape[0], Ho = (int)hw2_w->shape[0];
    IVLM_TRY(gemm(h, mt, C, hw0, C, tk_, Hh, nv, Hh, C, hb0, IVLM_ACT_RELU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream_));
    IVLM_TRY(gemm(h, tk_, Hh, hw1, Hh, tv, Hh, nv, Hh, Hh, hb1, IVLM_ACT_RELU, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream_));
    IVLM_TRY(gemm(h, tv, Hh, hw2, Hh, ta, Ho, nv, Ho, Hh, hb2, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream_));
    IVLM_W(up0_w, "dec.up0_w"); IVLM_W(up0_b, "dec.up0_b"); IVLM_W(up_lng, "dec.up_lng"); IVLM_W(up_lnb, "dec.up_lnb");
    IVLM_W(up3_w, "dec.up3_w"); IVLM_W(up3_b, "dec.up3_b");
    const int U = (int)up0_w_w->shape[0];   // 4 * co
    IVLM_REQUIRE(U == C && Ho == 32 && U / 4 == 64, "mask_decode: the fused upscaling covers SAM's 256 -> 64 -> 32 channels");
    IVLM_TRY(gemm(h, keys, C, up0_w, C, up_a, U, nv * S, U, C, up0_b, 0, nullptr, 0, nullptr, 0, -1, IVLM_BF16, stream_));
    IVLM_TRY(ivlm_layernorm_bf16(h, up_a, up_b, up_lng, up_lnb, (int64_t)nv * S * 4, U / 4, 1e-6f, nullptr, IVLM_ACT_GELU, stream_));
    return ivlm_upscale_hyper_dot(h, up_b, up3_w, up3_b, ta, a->lowres, nv, g, stream_);
}
