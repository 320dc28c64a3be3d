// Internal host-side runtime shared by the C-ABI translation units: context, error reporting,
// TMA tensor-map cache, bump workspace. Not part of the public ABI (see include/ivlm_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ivlm_b200.h"

namespace ivlm {

void set_error(const char* fmt, ...);

#define IVLM_CHECK_CUDA(expr)                                                                 \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ivlm::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,             \
                            cudaGetErrorString(_e));                                          \
            return IVLM_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define IVLM_REQUIRE(cond, ...)                                                               \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ivlm::set_error(__VA_ARGS__);                                                     \
            return IVLM_ERR_ARG;                                                              \
        }                                                                                     \
    } while (0)

#define IVLM_TRY(expr)                                                                        \
    do {                                                                                      \
        int _s = (expr);                                                                      \
        if (_s != IVLM_OK) return _s;                                                         \
    } while (0)

// Layout of the caller-provided workspace: [0, IVLM_WS_COUNTER_BYTES) int tile counters (zeroed by ivlm_set_workspace,
// self-resetting), then fp32 split-K partial tiles.
constexpr size_t IVLM_WS_COUNTER_BYTES = 65536;
constexpr size_t IVLM_TMAP_GEN = 2048;

struct TmapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    uint32_t box_cols = 64;
    uint32_t swizzle = 128;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols && swizzle == o.swizzle;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ k.rows;
        h = h * 1000003u ^ k.cols;
        h = h * 1000003u ^ k.ld;
        h = h * 1000003u ^ k.box_rows;
        h = h * 1000003u ^ (k.box_cols * 131u + k.swizzle);
        return h;
    }
};

struct Weight {
    const void* ptr = nullptr;
    int dtype = 0;  // IVLM_BF16 / IVLM_F32 / IVLM_I32
    int ndim = 0;
    int64_t shape[4] = {0, 0, 0, 0};
};

}  // namespace ivlm

struct ivlm_ctx {
    int device = 0;
    int num_sms = 148;
    uint64_t launches = 0;  // kernels launched through this handle (bench "gpu_launches")
    int sm_limit = 0;             // > 0: persistent token-major GEMMs launch at most this many CTAs (leaves SMs to a concurrent stream)
    int pdl = 0;                  // 1: launch the decode-chain kernels with programmatic dependent launch
    // Weight-streaming kernel for token counts <= 64 (gemv_small_m.cu).  Measured on B200 at 8 tokens (tools/prof_decode.py
    // sweep): for N = 5120 layers 16 rows x 10 warps per CTA gives o_proj 12.7 us / down_proj 30.8 us against 18.5 / 35.8
    // for the swapped tcgen05 kernel with fused split-K; for the wide layers (qkv 15360, gate-up 27648, lm_head) the
    // tcgen05 kernel is as fast or faster, so those keep it.
    int gv_rows8_max_n = 0;       // layers with N <= this use 8 rows per CTA (A/B knob; 16 rows measured better)
    int gv_warps = 10;            // warps per CTA (0: heuristic)
    int gv_max_n = 8192;          // layers wider than this go to the swapped tcgen05 kernel instead
    int fused_split_force = 0;    // > 0: fused split-K factor of the swapped GEMM instead of the cost model's choice (A/B knob)
    int gv_max_m = 8;             // token counts above this too: at 16 / 32 / 64 tokens the swapped tcgen05 kernel wins (o_proj 20 / 29 / 33 us
                                  // against 23 / 48 / 52, down_proj 40 / 57 / 62 against 56 / 136 / 130; tools/prof_decode.py variants)
    int small_m_variant = 0;      // 0: weight-streaming mma.sync kernel for token counts <= 64; 1: swapped-operand tcgen05 path
    int global_attn_variant = 0;  // 0: 64-key tiles, 2 CTAs/SM; 1: 128-key tiles, 1 CTA/SM (A/B switch)
    int attn_variant = 0;         // 1: plain / causal attention on the mma.sync flash kernel instead of the tcgen05 one (A/B, fallback)
    int attn_small_variant = 0;   // 1: per-query sweeps in the few-queries decoder attention (the round-1 kernel; A/B and bit-identity tests)
    int ds_prefetch_kb = 0;       // decode_stream L2 prefetch of the successor's weights per CTA: 0 off (default: measured 1-10 % SLOWER on
                                  // the 13B chain, profiles/r2_decode_layer_ops_in_graph.txt), -1 as the caller asks, > 0 cap in KB
    int ds_stages = 0;            // decode_stream ring depth (0: 6 stages -- 137 vs 139 us per layer with 8 in the same run; A/B knob)
    int ds_force_stream = 1;      // decode_stream: 1 streams the activation with the weights whenever no RMSNorm is fused (no prologue: o_proj
                                  // 12.8 vs 13.7 us), 0 keeps it resident when it fits
    int dec_prefetch = 0;         // paged decode attention: 1 requests the CTA's K / V lines into L2 before the page loops (A/B knob; measured
                                  // SLOWER, 19.6 vs 16.0 us per layer: the requests queue in front of the demand loads they were meant to hide)
    int dec_warps = 0;            // paged decode attention: warps per CTA (0: 8; 11 or 16 for A/B)
    int attn_prefetch_ahead = 0;  // window attention: L2 prefetch of the successor CTA's tiles, distance in CTAs of the launch order
                                  // (0 = off, the default: measured neutral at 148 .. 1184 CTAs ahead, 0.366-0.379 ms per 16 views)
    int window_attn_variant = 0;  // 0: single-tile 2-CTA/SM window kernel, 1: the general tiled kernel (A/B switch)
    // TMA descriptor cache in two generations: lookups see both, inserts go to `tmaps`; when it holds IVLM_TMAP_GEN entries it
    // becomes `tmaps_old` (whose previous content is dropped).  A pointer handed out therefore stays valid for at least
    // IVLM_TMAP_GEN further insertions -- one call fetches at most a handful -- while activations that churn addresses cannot
    // grow the cache without bound.  (unordered_map never moves its nodes on insert.)
    std::unordered_map<ivlm::TmapKey, CUtensorMap, ivlm::TmapKeyHash> tmaps, tmaps_old;
    uint64_t attr_done = 0;       // bit i: cudaFuncSetAttribute done on this handle's device for kernel variant i (per handle = per device)
    std::unordered_map<std::string, ivlm::Weight> weights;   // ivlm_bind_weights: borrowed device pointers by name
    ivlm_model_dims dims = {};                               // ivlm_set_model_dims
    int dims_set = 0;
    void* jpeg = nullptr;                                    // lazily created nvJPEG handle + state (jpeg.cu)
    // caller-provided scratch (bump allocated inside stage drivers)
    char* ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_off = 0;

    void* ws_alloc(size_t bytes) {
        size_t a = (ws_off + 255) & ~size_t(255);
        if (a + bytes > ws_bytes) return nullptr;
        ws_off = a + bytes;
        return ws + a;
    }
};

namespace ivlm {
// Kernel launch with (optionally) the programmatic-stream-serialization attribute.  Only for kernels that execute
// pdl_wait() before reading their inputs.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(const ivlm_ctx* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Returns (creating and caching if needed) a 2D bf16 tensor map: rows x cols, row pitch ld elements,
// box = box_rows x 64 elements, 128B swizzle, zero OOB fill.
int get_tmap_bf16(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  const CUtensorMap** out);
// Rank-3 (k-chunk, row, chunk index) map of a [rows, K] matrix, box (64, box_rows, 8), 128B swizzle (decode_stream.cu).
int get_tmap_bf16_kchunk3d(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows,
                           const CUtensorMap** out);
// Weight-streaming small-token GEMM (gemv_small_m.cu); arguments as ivlm_gemm_bf16.
int launch_gemv_small_m(ivlm_ctx* h, const ivlm_gemm_args* a, cudaStream_t stream);
// General form: box = box_rows x box_cols elements, swizzle_bytes in {128, 64, 32} (box_cols * 2 must not exceed it).
// plain / causal attention on tcgen05 (attention_tcgen05.cu): 1 = launched, 0 = layout not covered (use flash_attn_kernel), < 0 error
int attention_tcgen05_try(ivlm_ctx* h, const ivlm_attn_args* a, cudaStream_t stream);
int get_tmap_bf16_ex(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols, uint32_t swizzle_bytes, const CUtensorMap** out);
}  // namespace ivlm
