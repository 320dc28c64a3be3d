/* ivlm_b200 -- C ABI of the B200-native (sm_100a) InteractVLM inference hot path.
 *
 * The reference (saidwivedi/InteractVLM) has no FFI: its boundary is the Python class
 * InteractVLMForCausalLM (model/InteractVLM.py:139-638).  interactvlm_b200/model.py mirrors that
 * class and drives the kernels below through ctypes.  Every entry point names the reference code it
 * replaces.  Conventions:
 *   - all pointers are DEVICE pointers unless the name ends in _h (host);
 *   - bf16 tensors are raw uint16 storage, row-major, innermost dimension contiguous;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *     there is no hidden cudaDeviceSynchronize and no internal cudaMalloc on the hot path
 *     (ivlm_lift_build_* is the one-time exception and says so);
 *   - return 0 on success, negative ivlm_status otherwise; ivlm_last_error() gives the message;
 *   - a handle is bound to one device, not thread-safe; independent handles may run concurrently.
 */
#ifndef IVLM_B200_H
#define IVLM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IVLM_API __attribute__((visibility("default")))

typedef struct ivlm_ctx* ivlm_handle;

enum ivlm_status { IVLM_OK = 0, IVLM_ERR_ARG = -1, IVLM_ERR_CUDA = -2, IVLM_ERR_NOMEM = -3 };
enum ivlm_dtype { IVLM_BF16 = 0, IVLM_F32 = 1, IVLM_I32 = 2, IVLM_I64 = 3 };
enum ivlm_act { IVLM_ACT_NONE = 0, IVLM_ACT_GELU = 1, IVLM_ACT_QUICK_GELU = 2, IVLM_ACT_RELU = 3, IVLM_ACT_SILU = 4,
                /* ivlm_gemm_bf16 only (token count > 64, bf16 output, no bias / residual): `w` holds gate / up rows interleaved in
                 * blocks of 8 (IVLM_EPI_SWIGLU's order) and the epilogue writes out[m, f] = bf16(bf16(silu(gate)) * up), N / 2
                 * columns -- HF LlamaMLP's gate without the [M, 2F] round trip through HBM */
                IVLM_ACT_SWIGLU = 5 };
enum ivlm_lift_mode {
    IVLM_LIFT_HUMAN = 0,       /* HumanContact3DPredictor: clamp +-20, sigmoid, final clamp [0,1] */
    IVLM_LIFT_OBJECT_MESH = 1, /* ObjectMeshContact3DPredictor: sigmoid, only pixels with p > thr vote */
    IVLM_LIFT_POINTS = 2       /* ObjectPCAfford3DPredictor: raw values, unit weights */
};

IVLM_API int ivlm_create(ivlm_handle* out, int device);
IVLM_API int ivlm_destroy(ivlm_handle h);
IVLM_API const char* ivlm_last_error(void);
IVLM_API int ivlm_abi_version(void);
/* Caller-owned scratch in device memory (>= 1 MiB; 32 MiB covers every shape on the path).  With a workspace bound,
 * weight-streaming GEMMs (small token counts) split K across CTAs and reduce the partials in-kernel, deterministically. */
IVLM_API int ivlm_set_workspace(ivlm_handle h, void* ptr, size_t bytes, void* stream);
/* Tuning / A-B switches (integer options): "window_attn_variant" 0 = single-tile window kernel, one softmax thread per query
 * row (default), 2 = the same with two threads per row, 1 = tiled kernel; "global_attn_variant" 0 = 64-key tiles, two softmax
 * threads per query row, single-pass softmax, separate K / V rings, 2 CTAs/SM (default), 2 = round-1 kernel (one thread per
 * row), 1 = 128-key tiles / 1 CTA per SM; "small_m_variant" 0 = weight-streaming kernel for token counts <= 64 (default),
 * 1 = swapped-operand tcgen05 kernel with fused split-K;
 * "pdl" 1 = launch the LLaMA decode-chain kernels with programmatic dependent launch (prologues overlap the predecessor's
 * tail; every such kernel executes griddepcontrol.wait before reading its inputs);
 * "sm_limit" n > 0 = persistent token-major GEMMs launched through this handle use at most n CTAs (one per SM), leaving
 * the other SMs to a second handle/stream (SAM encoder next to the weight-streaming decode chain); 0 = all SMs. */
IVLM_API int ivlm_set_option(ivlm_handle h, const char* name, int32_t value);
/* kernels launched through this handle so far (bench.py's "gpu_launches") */
IVLM_API uint64_t ivlm_launch_count(ivlm_handle h);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (csrc/stages.cu): bind the checkpoint tensors once, then one call per stage.  What a reference
 * maintainer would bind instead of the Python-level loops of InteractVLMForCausalLM: `get_visual_embs` ->
 * ImageEncoderViT.forward (InteractVLM.py:251-261, image_encoder.py:110-125), `LlavaLlamaForCausalLM.forward` over the prompt
 * (llava_llama.py:55-135) and one greedy-search step (transformers generation/utils.py as driven by InteractVLM.py:524-531).
 * Each driver is the fixed launch sequence of the op-level entry points below (bit-identical results); activations live in a
 * caller-provided arena; every launch goes to `stream`; nothing synchronises or allocates. */
typedef struct ivlm_weight_desc {
    const char* name;   /* "sam.w_patch", "sam.blocks.7.wqkv", "llm.3.wgu", "llm.lm_head", ... (interactvlm_b200/model.py lists them) */
    const void* ptr;    /* device pointer, borrowed: the caller keeps the tensor alive */
    int32_t dtype;      /* ivlm_dtype */
    int32_t ndim;
    int64_t shape[4];
} ivlm_weight_desc;
IVLM_API int ivlm_bind_weights(ivlm_handle h, const ivlm_weight_desc* descs, int32_t n);
typedef struct ivlm_model_dims {
    int32_t sam_img, sam_patch, sam_embed_dim, sam_depth, sam_heads, sam_window, sam_out_chans;
    uint32_t sam_global_mask;      /* bit i set: block i uses global attention (build_sam.py global_attn_indexes) */
    int32_t llm_hidden, llm_intermediate, llm_layers, llm_heads, llm_head_dim, llm_vocab;
    float llm_rms_eps;
    int32_t llm_paired_layout;     /* 1: q/k rows paired and gate/up rows interleaved (interactvlm_b200/layout.py) */
    /* CLIP tower (0 = not described: ivlm_clip_encode unavailable) */
    int32_t clip_img, clip_patch, clip_hidden, clip_heads, clip_layers;
    int32_t clip_ldk;              /* row pitch of the flattened patch-conv weight / im2col operand (3*patch^2 rounded up to 8) */
    float clip_eps;
} ivlm_model_dims;
IVLM_API int ivlm_set_model_dims(ivlm_handle h, const ivlm_model_dims* dims);
/* SAM ViT encoder on N views: images [N,3,S,S] bf16 -> emb [N, (S/patch)^2, out_chans] bf16 (token-major, the layout the mask
 * decoder reads).  win_map [N*nw*nw*ws*ws]: window-major row -> token row or -1 (window_partition with zero padding,
 * image_encoder.py:263-288); win_inv [N*(S/patch)^2]: its inverse; win_pads [n_pads]: the window-major rows that are padding. */
typedef struct ivlm_sam_encode_args {
    const void* images;
    void* emb;
    int32_t N;
    const int32_t* win_map;
    const int32_t* win_inv;
    const int32_t* win_pads;
    int32_t n_pads;
    void* arena;
    size_t arena_bytes;
} ivlm_sam_encode_args;
IVLM_API size_t ivlm_sam_encode_arena_bytes(ivlm_handle h, int32_t N);
IVLM_API int ivlm_sam_encode(ivlm_handle h, const ivlm_sam_encode_args* args, void* stream);
/* LLaMA prefill over B sequences of S embedded rows (right-padded): K/V pages, normed hidden states, greedy next token. */
typedef struct ivlm_llm_prefill_args {
    const void* embeds;         /* [B*S, D] bf16 */
    const int32_t* positions;   /* [B*S] */
    const int32_t* slot_map;    /* [B*S] cache slot of every row */
    void* const* k_cache;       /* HOST array of n_layers device pointers, each [pages, H, page_size, hd] bf16 */
    void* const* v_cache;
    void* hidden;               /* [B, max_len, D] bf16: rows [0,S) of every sequence are written */
    int32_t* next_tok;          /* [B] */
    const int32_t* last_rows;   /* [B] flat row b*S + S_b - 1 of every sequence's last valid row, or NULL (row S-1) */
    int32_t B, S, max_len, page_size;
    void* arena;
    size_t arena_bytes;
} ivlm_llm_prefill_args;
IVLM_API size_t ivlm_llm_arena_bytes(ivlm_handle h, int32_t tokens);
IVLM_API int ivlm_llm_prefill(ivlm_handle h, const ivlm_llm_prefill_args* args, void* stream);
/* CLIP ViT tower (all but the last layer) + patch-feature selection + mm_projector: what LlavaMetaForCausalLM.encode_images runs
 * (clip_encoder.py:31-60, llava_arch.py:93-96).  images [B,3,S,S] bf16 -> feats [B, T-1, llm_hidden] bf16 (T = (S/patch)^2 + 1).
 * patch_rows [B*(T-1)]: b*T + 1 + j; cls_rows [B]: b*T.  Weights "clip.*", "mm.w", "mm.b". */
typedef struct ivlm_clip_encode_args {
    const void* images;
    void* feats;
    int32_t B;
    const int32_t* patch_rows;
    const int32_t* cls_rows;
    void* arena;
    size_t arena_bytes;
} ivlm_clip_encode_args;
IVLM_API size_t ivlm_clip_encode_arena_bytes(ivlm_handle h, int32_t B);
IVLM_API int ivlm_clip_encode(ivlm_handle h, const ivlm_clip_encode_args* args, void* stream);
/* text_hidden_fcs[0] on the n hidden rows that predict [SEG] + the per-view camera gate (InteractVLM.py:100-112,268-294,551-556;
 * components.py:541-572): hidden_rows [n, llm_hidden] bf16, cam [n,V,5] bf16 (NULL without camera conditioning) ->
 * prompt [n,V,256] bf16 and emb [n,256] bf16 (the un-gated embedding).  Weights "seg.fc0_w" ... "seg.cam.*" (vi_v1 or none; the
 * other encoder / token types stay on the op-level path).  arena >= n * 5120 * 2 bytes + 256. */
IVLM_API int ivlm_seg_head(ivlm_handle h, const void* hidden_rows, const void* cam, void* prompt, void* emb, int32_t n, int32_t V,
                           void* arena, size_t arena_bytes, void* stream);
/* Prompt encoder + two-way mask decoder + output upscaling + hypernetwork dot for n samples of V views (InteractVLM.py:40-63,
 * prompt_encoder.py:140-186, mask_decoder.py:116-164, transformer.py:62-242): emb [n*V, S, C] bf16 token-major (ivlm_sam_encode's
 * output), prompt [n,V,C] bf16 -> low-res logits [n*V, 4g, 4g] fp32.  tok_idx [n*V*(5+V)]: for token row r of view (s,v), the row
 * of the table [5 output tokens; n*V prompt rows] it is made of (t < 5: t, else 5 + s*V + (t-5)).  Weights "dec.*". */
typedef struct ivlm_mask_decode_args {
    const void* emb;
    const void* prompt;
    float* lowres;
    int32_t n, V, heads;
    const int32_t* tok_idx;
    void* arena;
    size_t arena_bytes;
} ivlm_mask_decode_args;
IVLM_API size_t ivlm_mask_decode_arena_bytes(ivlm_handle h, int32_t n, int32_t V);
IVLM_API int ivlm_mask_decode(ivlm_handle h, const ivlm_mask_decode_args* args, void* stream);
/* One decode step for B <= 8 sequences (the launch sequence a CUDA graph captures): bookkeeping buffers as ivlm_decode_prepare /
 * ivlm_decode_finish take them. */
typedef struct ivlm_llm_decode_args {
    int32_t* state; int32_t S; const int32_t* S_rows; const int32_t* scripted; int32_t G;
    int32_t* next; int32_t* done; int32_t* out_tokens; int32_t* tok; int32_t* pos; int32_t* slot; int32_t* seq_lens;
    const int32_t* slot_base; int32_t eos, pad, B;
    void* const* k_cache; void* const* v_cache;   /* HOST arrays of n_layers device pointers */
    const int32_t* block_table; int32_t max_pages, page_size;
    void* hidden; int32_t max_len;                /* [B, max_len, D] */
    void* hid_step;                               /* [B, D] */
    void* arena; size_t arena_bytes;              /* ivlm_llm_arena_bytes(h, B) */
} ivlm_llm_decode_args;
IVLM_API int ivlm_llm_decode_step(ivlm_handle h, const ivlm_llm_decode_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction (tcgen05 / TMEM / TMA).  Replaces every nn.Linear / 1x1-or-patch Conv2d call on
 * the path: SAM image_encoder.py:235-260 (qkv, proj), common.py:13-26 (MLPBlock), image_encoder.py:
 * 418-426 (PatchEmbed as im2col GEMM), :92-108 (neck); HF CLIPVisionModel / LlamaModel projections
 * (clip_encoder.py:46-56, llava_llama.py:93-105); mm_projector (llava_arch.py:93-96);
 * text_hidden_fcs (InteractVLM.py:100-112); mask-decoder projections (transformer.py:205-242).
 *   out[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + residual[M,N]
 * With out_dtype == IVLM_BF16 the epilogue rounds to bf16 after the bias add, after the activation
 * and after the residual add, like the eager bf16 reference does at those op boundaries.
 */
typedef struct ivlm_gemm_args {
    const void* a;        /* [M,K] bf16 */
    int64_t lda;
    const void* w;        /* [N,K] bf16 (nn.Linear.weight layout) */
    int64_t ldw;
    void* out;            /* [M,N] bf16 or fp32 */
    int64_t ldo;
    const void* bias;     /* [N] bf16 or NULL */
    const void* residual; /* [M,N] bf16 or NULL (added after the activation) */
    int64_t ldr;
    const int32_t* row_map; /* optional [M]: output (and residual) row for A-row m; -1 drops the row */
    int32_t M, N, K;
    int32_t act;          /* ivlm_act */
    int32_t out_dtype;    /* IVLM_BF16 or IVLM_F32 */
    int32_t k_splits;     /* >1: split-K, atomically accumulates raw fp32 into a pre-zeroed `out`; 0: let the library decide
                             (fused deterministic split-K when a workspace is bound); 1: never split */
    int32_t force_swap;   /* 0 auto, 1 weights-as-128-row-operand, -1 never */
    int32_t no_round;     /* 1: skip the intermediate bf16 roundings */
    int32_t res_row_mod;  /* >0: residual row = out_row % res_row_mod (broadcast tables, e.g. pos_embed) */
} ivlm_gemm_args;
IVLM_API int ivlm_gemm_bf16(ivlm_handle h, const ivlm_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise normalisation and elementwise kernels. */
/* y[m,:] = act(LayerNorm(x[row_map[m],:])) (zeros if row_map[m] < 0).  nn.LayerNorm / LayerNorm2d
 * (common.py:31-43) on token-major rows; row_map implements window_partition + zero pad
 * (image_encoder.py:263-288) when given. */
IVLM_API int ivlm_layernorm_bf16(ivlm_handle h, const void* x, void* y, const void* gamma, const void* beta,
                        int64_t out_rows, int32_t D, float eps, const int32_t* row_map, int32_t act, void* stream);
/* HF LlamaRMSNorm: y = gamma * bf16(x * rsqrt(mean(x^2) + eps)), variance in fp32. */
IVLM_API int ivlm_rmsnorm_bf16(ivlm_handle h, const void* x, void* y, const void* gamma, int64_t rows, int32_t D, float eps,
                      void* stream);
/* out[i] = bf16(a[i] + b[i % period]) (period == 0 -> n).  `keys + key_pe`, `x + pos_embed`. */
IVLM_API int ivlm_add_bcast_bf16(ivlm_handle h, const void* a, const void* b, void* out, int64_t n, int64_t period,
                        void* stream);
/* SwiGLU gate: out[r,j] = bf16(bf16(silu(gate[r,j])) * up[r,j]); HF LlamaMLP.  gate_up [rows, 2F]: gate = columns [0,F), up =
 * [F,2F); with interleaved != 0 gate/up alternate in blocks of 8 columns (the weight-row order of IVLM_EPI_SWIGLU). */
IVLM_API int ivlm_silu_mul_bf16(ivlm_handle h, const void* gate_up, void* out, int64_t rows, int32_t F, int32_t interleaved,
                       void* stream);
/* out[rows[i], 0:N] = vec[0:N] (bf16) for i < n_rows: the q/k/v rows of the SAM window padding are the qkv bias (the padded
 * tokens are zeros after norm1, image_encoder.py:179-183), so they are written as a broadcast instead of being multiplied. */
IVLM_API int ivlm_fill_rows_bf16(ivlm_handle h, void* out, int64_t ld, const int32_t* rows, int32_t n_rows, const void* vec, int32_t N,
                        void* stream);
/* fp32 split-K accumulator -> bf16 with optional bias / activation / residual. */
IVLM_API int ivlm_finalize_f32_bf16(ivlm_handle h, const float* acc, void* out, const void* bias, const void* residual,
                           int64_t rows, int32_t N, int32_t act, void* stream);
/* x [n] fp32 -> bf16 and back (utility for staging inputs). */
IVLM_API int ivlm_cast_f32_bf16(ivlm_handle h, const float* x, void* y, int64_t n, void* stream);
IVLM_API int ivlm_cast_bf16_f32(ivlm_handle h, const void* x, float* y, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Patch / conv lowering. */
/* img [N,C,H,W] bf16 -> cols [N*(H/p)*(W/p), ldk] with k = c*p*p + dy*p + dx (Conv2d weight flatten order),
 * columns [C*p*p, ldk) zero filled.  PatchEmbed (image_encoder.py:395-426), CLIP patch_embedding. */
IVLM_API int ivlm_im2col_patch_bf16(ivlm_handle h, const void* img, void* cols, int32_t N, int32_t C, int32_t H, int32_t W,
                           int32_t p, int32_t ldk, void* stream);
/* x [N,H,W,C] token-major bf16 -> cols [N*H*W, 9*C] with k = (ky*3+kx)*C + c, zero padding 1.
 * SAM neck 3x3 conv (image_encoder.py:99-106); the weight is repacked to [Cout,3,3,Cin] at bind time. */
IVLM_API int ivlm_im2col_3x3_bf16(ivlm_handle h, const void* x, void* cols, int32_t N, int32_t H, int32_t W, int32_t C,
                         void* stream);

/* Input pipeline (run_demo.py:65-79 `preprocess`, CLIPImageProcessor rescale + normalise): uint8 HWC images [N,H,W,3]
 * (already resized on the host like the reference does) -> bf16 CHW [N,3,S,S], (x*pre_scale - mean[c]) / std[c], zero
 * padded to S x S.  mean3_h / std3_h are HOST arrays of 3 floats. */
IVLM_API int ivlm_preprocess_u8_bf16(ivlm_handle h, const uint8_t* img, void* out, int32_t N, int32_t H, int32_t W, int32_t S,
                            float pre_scale, const float* mean3_h, const float* std3_h, void* stream);

/* JPEG decode through nvJPEG (a library call, as cuBLAS would be for a plain GEMM): the compressed file in HOST memory -> interleaved
 * RGB uint8 [H,W,3] in DEVICE memory, the input of ivlm_resample_u8 / ivlm_preprocess_u8_bf16.  Replaces the cv2.imread of
 * run_demo.py:330 / the dataset classes on the way to the GPU; pixel values may differ from libjpeg-turbo's by a few grey levels. */
IVLM_API int ivlm_jpeg_info(ivlm_handle h, const uint8_t* data_h, size_t n, int32_t* height, int32_t* width);
IVLM_API int ivlm_jpeg_decode_rgb(ivlm_handle h, const uint8_t* data_h, size_t n, uint8_t* rgb, int32_t H, int32_t W, void* stream);

/* One pass (vertical = 0: along x, 1: along y) of Pillow's antialiased 8-bit resize -- the arithmetic of the reference's
 * ResizeLongestSide.apply_image (segment_anything/utils/transforms.py:27-34, bilinear) and CLIPImageProcessor resize
 * (bicubic): src [N,H,W,3] uint8 -> dst [N,OH,OW,3]; bounds [out,2] (first input index, count) and coeffs [out,ksize]
 * (22-bit fixed point) are DEVICE arrays built by the host (interactvlm_b200/resample.py).  Bit-exact vs Pillow. */
IVLM_API int ivlm_resample_u8(ivlm_handle h, const uint8_t* src, uint8_t* dst, const int32_t* bounds, const int32_t* coeffs,
                     int32_t ksize, int32_t N, int32_t H, int32_t W, int32_t OH, int32_t OW, int32_t vertical, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention. */
typedef struct ivlm_attn_args {
    const void* q; const void* k; const void* v; void* out;  /* bf16 */
    int64_t q_bs, q_ts, q_hs;   /* element strides: batch, token, head (head_dim contiguous) */
    int64_t k_bs, k_ts, k_hs;
    int64_t v_bs, v_ts, v_hs;
    int64_t o_bs, o_ts, o_hs;
    int32_t B, H, Sq, Sk, D;    /* D (head_dim) in {64, 80, 128} */
    float scale;
    int32_t causal;             /* query i attends keys j <= i + (Sk - Sq) */
    const float* rel_h;         /* optional decomposed rel-pos bias [B,H,Sq,kh] (fp32) */
    const float* rel_w;         /* [B,H,Sq,kw]; key j -> (j / kw, j % kw) */
    int32_t kh, kw;
} ivlm_attn_args;
/* Fused softmax(QK^T*scale + bias)V.  SAM Attention.forward (image_encoder.py:235-260), HF CLIPAttention,
 * HF LlamaAttention prefill (eager path of transformers 4.31). */
IVLM_API int ivlm_attention_bf16(ivlm_handle h, const ivlm_attn_args* args, void* stream);
/* SAM decomposed relative position terms (image_encoder.py:354-392): rel_h[b,h,q,kh] = q . Rh[qy-kh+H-1],
 * rel_w[b,h,q,kw] = q . Rw[qx-kw+W-1]; q read from the packed qkv rows [B*S, 3*heads*hd]. */
IVLM_API int ivlm_sam_relpos(ivlm_handle h, const void* qkv, const void* rel_pos_h, const void* rel_pos_w, float* rel_h,
                    float* rel_w, int32_t B, int32_t heads, int32_t Hq, int32_t Wq, int32_t hd, void* stream);
/* SAM ViT attention with the decomposed relative-position bias computed in-kernel (image_encoder.py:235-260, :354-392)
 * on the tcgen05 tensor cores: qkv [B*S, 3*heads*80] packed rows (S = Hq*Wq tokens per image or window),
 * rel_pos_h [2*Hq-1, 80], rel_pos_w [2*Wq-1, 80] bf16 -> out rows [B*S] with pitch out_ld, columns head*80 + c.
 * Token grids 64x64 (global blocks) and 14x14 (windows).  out_row_map (optional, 14x14 windows only, [B*S] int32): output row of
 * input row r, -1 drops it -- window_unpartition (image_encoder.py:291-318) fused into the store, so the rows of the 64 -> 70
 * zero padding are never written and the proj GEMM that follows runs on the real tokens only. */
IVLM_API int ivlm_sam_attention_bf16(ivlm_handle h, const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out,
                            int32_t B, int32_t heads, int32_t Hq, int32_t Wq, int32_t hd, int64_t out_ld,
                            const int32_t* out_row_map, void* stream);
/* Attention with few queries or few keys and small head_dim (SAM TwoWayTransformer, transformer.py:185-242):
 * q [B,Nq,heads*hd], k/v [B,Nk,heads*hd], out [B,Nq,heads*hd]; hd in {16,32}. q batch may be 1 (broadcast). */
IVLM_API int ivlm_attn_small_bf16(ivlm_handle h, const void* q, const void* k, const void* v, void* out, int32_t B,
                         int32_t q_batch_stride_zero, int32_t Nq, int32_t Nk, int32_t heads, int32_t hd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LLaVA / LLaMA glue. */
/* prepare_inputs_labels_for_multimodal (llava_arch.py:98-347, branch :185-208): embed_tokens gather with the
 * single IMAGE_TOKEN_INDEX (-200) of each row replaced by `n_img` projected CLIP rows.
 * ids [B,L] int32; out [B, L-1+n_img, D]. */
IVLM_API int ivlm_embed_splice_bf16(ivlm_handle h, const void* embed, const int32_t* ids, const void* img_feats, void* out,
                           int32_t B, int32_t L, int32_t n_img, int32_t D, int32_t vocab, void* stream);
/* embed_tokens gather for decode steps: out[b,:] = embed[ids[b],:] */
IVLM_API int ivlm_embed_gather_bf16(ivlm_handle h, const void* embed, const int32_t* ids, void* out, int32_t n, int32_t D,
                           int32_t vocab, void* stream);
/* HF apply_rotary_pos_emb (rotate_half, bf16 cos/sin tables [max_pos, hd]) on the packed qkv rows
 * [T, 3*H*hd]; writes rotated q [T,H*hd], and k/v both contiguous [T,H*hd] (may be NULL) and into the paged
 * KV cache at slot_map[t] = physical_page * page_size + offset (cache layout [pages, H, page_size, hd]: the tokens
 * of one head inside a page are contiguous).  paired != 0: the q and k COLUMNS of qkv are in the paired order of
 * IVLM_EPI_ROPE_KV (feature j of a head at (j/8)*16 + j%8, feature j + hd/2 eight columns later); outputs are always natural. */
IVLM_API int ivlm_rope_kv_store_bf16(ivlm_handle h, const void* qkv, const int32_t* positions, const int32_t* slot_map,
                            const void* cos_t, const void* sin_t, void* q_out, void* k_out, void* v_out,
                            void* k_cache, void* v_cache, int32_t T, int32_t H, int32_t hd, int32_t page_size,
                            int32_t paired, void* stream);
/* Decode-loop bookkeeping on the device (HF greedy search, transformers 4.31 generation/utils.py, as driven by
 * InteractVLM.py:524-531), so that one decode step is a pure CUDA-graph replay:
 * prepare: step = state[0]++ (also copied to state[1]); token fed at this step = scripted[b,step] or next[b]
 * (pad_token once the sample has emitted eos); records it in out_tokens[b,step]; sets tok/pos/slot/seq_lens for
 * position S_b+step, S_b = S_rows[b] (per-sample prompt rows of a right-padded batch, llava_arch.py:98-347 pads the same way)
 * or the scalar S when S_rows == NULL.  finish: hidden[b, S_b+step, :] = hid_step[b, :]. */
IVLM_API int ivlm_decode_prepare(ivlm_handle h, int32_t* state, int32_t S, const int32_t* S_rows, const int32_t* scripted, int32_t G,
                        const int32_t* next, int32_t* done, int32_t* out_tokens, int32_t* tok, int32_t* pos, int32_t* slot,
                        int32_t* seq_lens, const int32_t* slot_base, int32_t eos, int32_t pad, int32_t B, void* stream);
IVLM_API int ivlm_decode_finish(ivlm_handle h, const int32_t* state, int32_t S, const int32_t* S_rows, const void* hid_step,
                       void* hidden, int32_t B, int32_t D, int32_t max_len, void* stream);
/* Weight-streaming linear layer of a LLaMA decode step (token count M <= 8) with the row-wise neighbours fused in
 * (csrc/decode_stream.cu): optional HF LlamaRMSNorm prologue on the activation rows, and one of three epilogues --
 *   IVLM_EPI_PLAIN    out = act(A W^T + bias) + residual (o_proj, down_proj, lm_head of HF LlamaDecoderLayer / lm_head);
 *   IVLM_EPI_SWIGLU   HF LlamaMLP: out[m,f] = bf16(bf16(silu(gate[m,f])) * up[m,f]); `w` holds gate/up rows INTERLEAVED in blocks
 *                     of 8 (rows 16t..16t+7 = gate features 8t..8t+7, rows 16t+8..16t+15 = up features 8t..8t+7); out [M, N/2];
 *   IVLM_EPI_ROPE_KV  HF apply_rotary_pos_emb + cache write: `w` = [q;k;v] rows with q/k rows PAIRED per head in blocks of 8
 *                     (rows h*hd+16b..+7 = features h*hd+8b..+7, rows +8..+15 = the same features + hd/2), v rows natural;
 *                     rotated q -> out [M, H*hd] (natural order), rotated k and v -> paged cache at slot_map[m]
 *                     (layout of ivlm_rope_kv_store_bf16).
 * Replaces, per decoder layer, the separate rmsnorm / rope_kv_store / silu_mul launches of the chain (HF 4.31
 * modeling_llama.py LlamaDecoderLayer.forward as driven by /root/reference model/llava/model/language_model/llava_llama.py:93-105).  Needs a bound workspace;
 * K must be a multiple of 64 and all operands 16-byte aligned (the weights arrive through a rank-3 TMA tensor map). */
enum ivlm_decode_epilogue { IVLM_EPI_PLAIN = 0, IVLM_EPI_SWIGLU = 1, IVLM_EPI_ROPE_KV = 2 };
typedef struct ivlm_decode_linear_args {
    const void* a;          /* [M,K] bf16 activations (the un-normalised residual stream when norm_gamma != NULL) */
    int64_t lda;
    const void* w;          /* [N,K] bf16, row order as the epilogue requires */
    int64_t ldw;
    int32_t M, N, K;
    const void* norm_gamma; /* [K] bf16 RMSNorm weight or NULL */
    float norm_eps;
    int32_t epilogue;       /* ivlm_decode_epilogue */
    int32_t act;            /* PLAIN only: ivlm_act */
    const void* bias;       /* PLAIN only: [N] bf16 or NULL */
    const void* residual;   /* PLAIN only: [M,N] bf16 or NULL */
    int64_t ldr;
    void* out;              /* PLAIN [M,N]; SWIGLU [M,N/2]; ROPE_KV q [M,H*hd] */
    int64_t ldo;
    int32_t out_dtype;      /* IVLM_BF16 (IVLM_F32 allowed for PLAIN: raw accumulators, e.g. lm_head logits) */
    /* ROPE_KV only */
    const int32_t* positions; /* [M] */
    const int32_t* slot_map;  /* [M] physical_page * page_size + offset */
    const void* cos_t;        /* [max_pos, hd] bf16 */
    const void* sin_t;
    void* k_cache;            /* [pages, H, page_size, hd] bf16 */
    void* v_cache;
    int32_t H, hd, page_size;
    /* optional: weights [prefetch_N, prefetch_K] of the NEXT ivlm_decode_linear launch on this stream.  Each CTA asks L2
     * (cp.async.bulk.prefetch.L2) for the first prefetch_stages 16 KB stages its successor CTA will stream, once its own last
     * stage is in flight -- the HBM pipe then stays busy across the launch boundary (and across the attention launch between
     * qkv and o_proj).  0 stages = 256 KB per SM.  OFF unless option "ds_prefetch_kb" is set (-1: as asked, > 0: cap in KB):
     * on the 13B chain it measured 1-10 % slower than no prefetch (the launches are bounded by their fixed costs, not by an
     * idle HBM pipe), so it stays an A/B knob.  Results do not depend on it. */
    const void* prefetch_w;
    int64_t prefetch_ldw;
    int32_t prefetch_N, prefetch_K, prefetch_stages;
} ivlm_decode_linear_args;
IVLM_API int ivlm_decode_linear(ivlm_handle h, const ivlm_decode_linear_args* args, void* stream);
/* n (1..4) ivlm_decode_linear launches that feed each other (phase i+1 reads what phase i wrote: o_proj -> gate/up ->
 * down_proj -> the next layer's qkv) as ONE launch of a persistent kernel: the phases are separated by grid-wide barriers and the
 * weight stream of phase i+1 is already in flight while phase i finishes, instead of paying a launch boundary per layer
 * (HF LlamaDecoderLayer.forward, the part after the attention; /root/reference model/llava/model/language_model/llava_llama.py:93-105
 * drives it).  Results are bit-identical to the separate launches.  Phases with norm_gamma keep their activation in shared
 * memory, the others stream it; at most one ROPE_KV phase; prefetch_* fields are ignored. */
IVLM_API int ivlm_decode_chain(ivlm_handle h, const ivlm_decode_linear_args* phases, int32_t n, void* stream);
/* One-token attention over the paged KV cache: q [B,H*hd], block_table [B,max_pages], seq_lens [B]
 * (keys 0..seq_len-1, the current token already stored); caches laid out [pages, H, page_size, hd]. */
IVLM_API int ivlm_decode_attention_paged_bf16(ivlm_handle h, const void* q, const void* k_cache, const void* v_cache,
                                     const int32_t* block_table, const int32_t* seq_lens, void* out, int32_t B,
                                     int32_t H, int32_t hd, int32_t page_size, int32_t max_pages, float scale,
                                     void* stream);
/* greedy token: argmax over fp32 logits [B, ld] (first `vocab` columns), lowest index wins ties (torch.argmax). */
IVLM_API int ivlm_argmax_f32(ivlm_handle h, const float* logits, int32_t* out, int32_t B, int32_t vocab, int64_t ld,
                    void* stream);
/* rows gather: out[i,:] = x[idx[i],:] (bf16), e.g. the [SEG]-1 hidden rows (InteractVLM.py:545-556). */
/* neq[i*K + k] = 1 when row i of x [n rows] and row k of ref [K rows] differ in any bit, else 0 (rows of row_bytes bytes,
 * a multiple of 16; 16-byte aligned).  Exact-match test of the encoder's view cache: the hcontact harness feeds the same
 * four body renders with every image (run_demo.py:279-281), whose embeddings the reference recomputes each time
 * (InteractVLM.py:578). */
IVLM_API int ivlm_rows_differ(ivlm_handle h, const void* x, int32_t n, const void* ref, int32_t K, int64_t row_bytes,
                     int32_t* neq, void* stream);
IVLM_API int ivlm_gather_rows_bf16(ivlm_handle h, const void* x, const int32_t* idx, void* out, int32_t n, int32_t D,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * Prompt side. */
/* VIv1CamPoseEncoder (components.py:541-572) + the gate multiply of process_embeddings
 * (InteractVLM.py:268-282): out[b,v,:] = bf16(emb[b,:] * sigmoid(W_v relu(W2 relu(W1 cam[b,v] + b1) + b2) + b_v)).
 * cam [B,V,5] bf16, emb [B,256] bf16, w1 [128,5], w2 [128,128], wv [V,256,128]. */
IVLM_API int ivlm_cam_gate_bf16(ivlm_handle h, const void* cam, const void* emb, const void* w1, const void* b1,
                       const void* w2, const void* b2, const void* wv, const void* bv, void* out, int32_t B, int32_t V,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Mask decoder tail. */
/* Second ConvTranspose2d(64->32,k2,s2)+GELU of output_upscaling fused with the hypernetwork dot product
 * (mask_decoder.py:143-157): up1 [Bv,64*64,4,64] bf16 (token, first-stage sub-pixel, channel),
 * w2 [4,32,64] bf16 (sub-pixel, out-ch, in-ch), b2 [32], hyper [Bv,32] bf16 -> low-res logits [Bv,256,256] fp32
 * holding bf16-rounded values (the reference emits bf16 masks then .float()s them, sam.py:161). */
IVLM_API int ivlm_upscale_hyper_dot(ivlm_handle h, const void* up1, const void* w2, const void* b2, const void* hyper,
                           float* lowres, int32_t Bv, int32_t grid, void* stream);
/* F.interpolate(mode="bilinear", align_corners=False) used twice by Sam.postprocess_masks (sam.py:137-172):
 * src [N,sh,sw] fp32, only the top-left (crop_h,crop_w) window is read -> dst [N,dh,dw]. */
IVLM_API int ivlm_bilinear_f32(ivlm_handle h, const float* src, float* dst, int32_t N, int32_t sh, int32_t sw, int32_t crop_h,
                      int32_t crop_w, int32_t dh, int32_t dw, void* stream);

/* x[i] = sigmoid(x[i]) where gt == NULL or gt[i] != ignore_value (InteractVLM.py:452-456, 'oafford' with HM view types). */
IVLM_API int ivlm_sigmoid_where_f32(ivlm_handle h, float* x, const float* gt, float ignore_value, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Render-Localise-Lift: 2D masks -> per-vertex contact. */
typedef struct ivlm_lift_map ivlm_lift_map;
/* One-time build (allocates device memory, synchronises): per-vertex CSR of (pixel, barycentric weight)
 * from the reference's maps pixel_to_vertex [V,H,W,3] int64 (-1 background) and bary [V,H,W,3] fp32 host
 * arrays (components.py:203-218; lift2d_dict.pkl, components.py:392-424).  A pixel votes only when all three
 * vertex ids are in [0, n_verts) (components.py:241-245). */
IVLM_API int ivlm_lift_build_mesh(ivlm_handle h, const int64_t* p2v_h, const float* bary_h, int32_t V, int32_t H, int32_t W,
                         int32_t n_verts, ivlm_lift_map** out);
/* Point-cloud variant: pixel_to_point [V,H,W] int64 (-1 background), unit weights (components.py:319-347). */
IVLM_API int ivlm_lift_build_points(ivlm_handle h, const int64_t* p2p_h, int32_t V, int32_t H, int32_t W, int32_t n_points,
                           ivlm_lift_map** out);
IVLM_API int ivlm_lift_free(ivlm_lift_map* m);
IVLM_API int64_t ivlm_lift_nnz(const ivlm_lift_map* m);
/* masks [B,V,H,W] fp32 logits (POINTS mode: values) -> contact [B,n_verts] fp32.
 * HumanContact3DPredictor.forward (components.py:220-277), ObjectMeshContact3DPredictor (components.py:430-489,
 * thr = 0.3), ObjectPCAfford3DPredictor (components.py:289-347). Deterministic gather, no atomics: one warp per (vertex,
 * view) strides the CSR segment with coalesced loads and folds the lanes with shuffles. */
IVLM_API int ivlm_lift(ivlm_handle h, const ivlm_lift_map* m, const float* masks, float* contact, int32_t B, int32_t mode,
              float thr, void* stream);
/* Same lift fed with the mask decoder's LOW-RES logits [B,V,sh,sw] fp32 (mask_decoder.py:116-164 output): the bilinear
 * upsample to the map's (H,W) that Sam.postprocess_masks applies first (sam.py:161-165, align_corners=False, no crop) is
 * evaluated per map entry inside the kernel with the arithmetic of ivlm_bilinear_f32, so the result is bit-identical to
 * ivlm_bilinear_f32 followed by ivlm_lift while reading 1/16 of the logits. */
IVLM_API int ivlm_lift_lowres(ivlm_handle h, const ivlm_lift_map* m, const float* lowres, int32_t sh, int32_t sw, float* contact,
                     int32_t B, int32_t mode, float thr, void* stream);
/* ------------------------------------------------------------------------------------------------
 * "Render" of Render-Localise-Lift: mesh -> per-pixel (face, barycentrics) -> lift maps and SAM input views.
 * Replaces pytorch3d's MeshRasterizer / HardPhongShader as the reference drives them (get_rasterizer,
 * project_vertices_and_create_mask, render_mesh: preprocess_data/render_mesh_utils.py:115-198; generate_sam_inp_objs:
 * utils/demo_utils.py:171-256): FoV perspective camera, blur_radius 0, faces_per_pixel 1, perspective-correct
 * barycentrics, z_clip = znear/2.  One-time preprocessing calls: they use the stream-ordered allocator for scratch and
 * synchronise `stream` once. */
#define IVLM_RASTER_MAX_VIEWS 8
typedef struct ivlm_raster_cam {
    float R[9];   /* row-major world->view rotation, row-vector convention: X_view = X_world R + T (look_at_view_transform) */
    float T[3];
    float C[3];   /* camera centre in world space (specular term of the shader) */
    float fx, fy; /* NDC focal lengths: x_ndc = fx * x_view / z_view + cx (FoV cameras: fx = fy = 1 / tan(fov / 2)) */
    float cx, cy; /* NDC principal point (0 for FoV cameras) */
    float z_clip; /* znear / 2; <= 0 disables clipping (PerspectiveCameras have no znear) */
} ivlm_raster_cam;
/* verts [Nv,3] fp32, faces [Nf,3] int32 (device) -> pix_to_face [V,H,W] int32 (-1 background), bary [V,H,W,3] fp32 (-1
 * background), optional zbuf [V,H,W] fp32 (-1 background) and pixel_to_vertices_map p2v [V,H,W,3] int64 (-1 background;
 * render_mesh_utils.py:140-163).  NDC +X left / +Y up (the longer image side spans [-L/S, L/S], the shorter [-1, 1]), pixel
 * centres, nearest depth wins, ties to the lower face index.
 * Faces entirely behind z_clip are culled, faces crossing it are kept where the interpolated depth is >= z_clip; faces
 * with a vertex at z <= 0 that are not culled are skipped and counted in *n_skipped_h (host, optional). */
IVLM_API int ivlm_rasterize_mesh(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                        const ivlm_raster_cam* cams_h, int32_t V, int32_t H, int32_t W, int32_t* pix_to_face, float* bary,
                        float* zbuf, int64_t* p2v, int32_t* n_skipped_h, void* stream);
/* pytorch3d PointsRasterizer as preprocess_data/utils_obj_pc.py:28-42,88-113 uses it for the LEMON / PIAD point clouds
 * (`num_point2pixel == 1`): points [n,3] fp32 (device) -> pixel_to_point map p2p [V,H,W] int64 = index of the point nearest in
 * depth among those within `radius` (NDC units) of the pixel centre, -1 where there is none.  Same NDC / pixel-centre
 * conventions as ivlm_rasterize_mesh; equal depths resolve to the lower index.  One-time preprocessing (stream-ordered scratch). */
IVLM_API int ivlm_rasterize_points(ivlm_handle h, const float* points, int32_t n_points, const ivlm_raster_cam* cams_h, int32_t V,
                          int32_t H, int32_t W, float radius, int64_t* p2p, void* stream);
/* HardPhongShader with one point light per view (lights_h [V,3] host) and default materials over the rasteriser output:
 * rgb [V,H,W,3] uint8 = trunc(255 * ((ambient + diffuse * max(n.l,0)) * colour + specular * max(v.r,0)^shininess)),
 * white background (render_mesh_utils.py:177-198).  colors [Nv,3] fp32 vertex colours (TexturesVertex). */
IVLM_API int ivlm_shade_phong(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                     const float* colors, const ivlm_raster_cam* cams_h, const float* lights_h, int32_t V, int32_t H, int32_t W,
                     const int32_t* pix_to_face, const float* bary, float ambient, float diffuse, float specular,
                     float shininess, uint8_t* rgb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pose refinement (optim/), contact term: ObjPose_Opt.contact_loss (optim/optimizer.py:80-96),
 *   loss = sum_ij p_i q_j |o_i - h_j| / (sum p * sum q)
 * and, when grad_obj != NULL, d loss / d obj_verts [n_obj,3] in the same pass (the reference builds the [n_obj, n_hum]
 * cdist and outer-product matrices and differentiates through them).  fp32; deterministic; uses the handle's workspace
 * (16 bytes x n_obj x splits). */
IVLM_API int ivlm_contact_loss(ivlm_handle h, const float* obj_verts, const float* obj_prob, const float* hum_verts,
                      const float* hum_prob, int32_t n_obj, int32_t n_hum, float* loss, float* grad_obj, void* stream);

/* Differentiable soft silhouette of the pose refinement: SSRenderer.render (optim/renderer.py:64-104) = pytorch3d
 * MeshRasterizer(blur_radius, faces_per_pixel = K, perspective-correct, clipped barycentrics) + SoftSilhouetteShader(sigma), one
 * camera.  Forward: alpha [H,W] = 1 - prod_k (1 - sigmoid(-d_k / sigma)) over the K fragments nearest in depth among the faces
 * within sqrt(blur_radius) of the pixel centre (signed squared NDC distance d_k, negative inside), zbuf0 [H,W] = depth of the
 * nearest fragment (-1 if none); the fragment lists n_frag [H,W], frag_face / frag_sd / frag_z [K,H,W] are kept for the
 * backward pass.  Backward: grad_verts [n_verts,3] = d(sum(grad_alpha * alpha)) / d verts, through the distance to the closest
 * edge and the projection (float atomics on the per-vertex NDC gradient).  One-time scratch from the stream-ordered allocator;
 * the forward synchronises `stream` once. */
IVLM_API int ivlm_soft_silhouette(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                         const ivlm_raster_cam* cam_h, int32_t H, int32_t W, float sigma, float blur_radius, int32_t K, float* alpha,
                         float* zbuf0, int32_t* n_frag, int32_t* frag_face, float* frag_sd, float* frag_z, void* stream);
IVLM_API int ivlm_soft_silhouette_backward(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                                  const ivlm_raster_cam* cam_h, int32_t H, int32_t W, float sigma, const float* grad_alpha,
                                  const int32_t* n_frag, const int32_t* frag_face, const float* frag_sd, float* grad_verts,
                                  void* stream);

/* Nearest neighbour (K = 1) of each of the n rows of x [n,D] among the m rows of y [m,D], 1 <= D <= 8, squared Euclidean
 * distance, ties to the lowest index: the knn_points(K=1) call of the contact ICP (optim/icp/icp.py:187-196, points ++ normals,
 * D = 6).  idx [n] int32, dist2 [n] fp32 (optional). */
IVLM_API int ivlm_knn1(ivlm_handle h, const float* x, const float* y, int32_t n, int32_t m, int32_t D, int32_t* idx, float* dist2,
              void* stream);

/* convert_contacts (utils/utils.py:428-443): SMPL->SMPL-X dense [R,C] matrix applied as CSR SpMV.
 * csr built once from the host dense matrix. */
typedef struct ivlm_csr ivlm_csr;
IVLM_API int ivlm_csr_build_dense(ivlm_handle h, const float* dense_h, int32_t rows, int32_t cols, ivlm_csr** out);
IVLM_API int ivlm_csr_free(ivlm_csr* m);
IVLM_API int ivlm_csr_spmv(ivlm_handle h, const ivlm_csr* m, const float* x, float* y, int32_t B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVLM_B200_H */
