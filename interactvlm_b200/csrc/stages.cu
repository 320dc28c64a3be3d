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
        IVLM_TRY(gemm(h, y, D, wgu, D, gu, 2 * F, T, 2 * F, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, IVLM_BF16, stream));
        IVLM_TRY(ivlm_silu_mul_bf16(h, gu, act, T, F, d.llm_paired_layout, stream));
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
