// Weight-streaming linear layer for LLaMA decode steps (token count M <= 8) with the neighbouring row-wise work fused in:
//
//   out[M,N] = epilogue( rmsnorm?(A)[M,K] . W[N,K]^T )
//
//   prologue (optional)  HF LlamaRMSNorm of the activation rows (transformers 4.31 modeling_llama.py: variance in fp32, value
//                        cast to bf16, times the bf16 weight) -- every CTA normalises the M x K activation into shared memory
//                        itself (80 KB read from L2), so the separate rmsnorm launch and its round trip disappear;
//   epilogue PLAIN       bias / activation / residual with the rounding points of the other GEMM kernels (o_proj, down_proj,
//                        lm_head);
//            SWIGLU      HF LlamaMLP gate: out[m,f] = bf16(bf16(silu(gate[m,f])) * up[m,f]); the weight rows are stored
//                        INTERLEAVED (8 gate rows, then the 8 up rows of the same features) so one 16-row tile holds both;
//            ROPE_KV     HF apply_rotary_pos_emb on q and k (bf16 products and sum, bf16 cos/sin tables) + the write of k, v
//                        into the paged cache and of q into its buffer; q/k weight rows are stored PAIRED (8 rows j, then
//                        the 8 rows j + hd/2 of the same head) so the two halves of a rotation meet in one tile.
//
// A decode step reads every weight once for <= 8 tokens: the launch is HBM-bound and built around the byte stream.
//   * persistent: one CTA per SM; the (16-row tile) x (512-column stage) grid of the weight matrix is cut into equal CONTIGUOUS
//     stage ranges, one per CTA (no wave quantisation: 320 or 960 tiles never divide by 148 SMs);
//   * one producer thread per CTA keeps an 8-deep ring of 16 KB stages full, ONE TMA op per stage: the weight matrix is described
//     as a 3-D tensor (64-element k-chunk, row, chunk index) and a (64, 16, 8) box lands as eight 128B-swizzled [16 rows x 128 B]
//     slabs -- the K-major layout ldmatrix reads without bank conflicts; >= 128 KB in flight per SM whatever the consumers do
//     (row-wise 1 KB cp.async.bulk copies, 16 per stage, measured 2.1 TB/s: the per-op issue cost dominates); the first ring of
//     weights is requested BEFORE griddepcontrol.wait (weights are static), so under programmatic dependent launch it streams
//     under the predecessor's tail;
//   * eight consumer warps split the 32 k16-steps of a stage (ldmatrix + mma.sync m16n8k16: 16 weight rows x 8 tokens), keep
//     their partial tile in registers across the stages of a tile and fold it through shared memory in warp order;
//   * a tile cut by a range boundary is finished by the CTA that holds its first stages: the CTA holding the rest computes that
//     part FIRST, parks it in the workspace and raises a flag (self-resetting; fixed summation order -> deterministic).
// No tensor-core tile shape drives this kernel (tcgen05 needs 128-row operands and buys nothing at 16 FLOP per byte).
#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int DS_ROWS = 16, DS_KW = 512, DS_STAGES = 8, DS_CONSUMERS = 8, DS_MAX_M = 8;
constexpr int DS_W_BYTES = DS_ROWS * DS_KW * 2, DS_A_BYTES = DS_MAX_M * DS_KW * 2;  // per stage: weights 16 KB, streamed tokens 8 KB
constexpr int DS_THREADS = (DS_CONSUMERS + 1) * 32;
constexpr int DS_BAR_BYTES = 256, DS_ROPE_HALF = 64;   // rotary tables of the ROPE_KV epilogue: head_dim <= 128
constexpr int DS_FIXED_BYTES = DS_BAR_BYTES + (DS_CONSUMERS + 1) * 128 * 4 + 2 * DS_MAX_M * DS_ROPE_HALF * 4 + 2 * DS_MAX_M * 4;
constexpr int DS_FLAG_OFFSET_BYTES = (int)IVLM_WS_COUNTER_BYTES - 4096;  // last 1024 ints of the counter region

enum { DS_EPI_PLAIN = 0, DS_EPI_SWIGLU = 1, DS_EPI_ROPE_KV = 2 };

struct DecodeStreamParams {
    const bf16* a;  long long lda;   // [M,K] activations (un-normalised when gamma != nullptr)
    const bf16* w;  long long ldw;   // [N,K] weights (row layout as the epilogue requires)
    int M, N, K;
    const bf16* gamma; float eps;    // fused RMSNorm or nullptr
    int epi, act, out_f32, round_steps;
    const bf16* bias; const bf16* res; long long ldr;
    void* out; long long ldo;
    // ROPE_KV
    const int* positions; const int* slot_map; const bf16* cos_t; const bf16* sin_t;
    bf16* k_cache; bf16* v_cache; int H, hd, page;
    // decomposition
    int tiles, spt, total_stages;    // spt = stages per tile
    int resident;                    // 1: the whole activation sits in shared memory (required with gamma)
    int nstages;                     // ring depth in use (<= DS_STAGES; option "ds_stages")
    float* partial; int* flags;
    // L2 prefetch of the SUCCESSOR's weights (the next launch of the decode chain): the first pf_n stages of the range its CTA
    // with this index will stream, requested once this CTA's own last stage is on its way
    const uint8_t* pf_w; long long pf_pitch; int pf_N, pf_rowbytes, pf_spt, pf_total, pf_grid, pf_n;
};

IVLM_DEVINL void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
IVLM_DEVINL void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
IVLM_DEVINL void ldmatrix_x2(uint32_t& r0, uint32_t& r1, const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(smem_row)));
}
IVLM_DEVINL int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
IVLM_DEVINL void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <bool RESIDENT>
__global__ void __launch_bounds__(DS_THREADS, 1)
decode_stream_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const DecodeStreamParams p) {
    extern __shared__ __align__(128) uint8_t ds_smem[];
    // [full bars | empty bars | rope bar] [red: 8 warps x 16 x 8 fp32] [fin: 16 x 8 fp32] [rope: cos, sin [8][64] fp32, page / offset
    // [8] int] [resident activation] [ring, 1024-byte aligned]
    uint64_t* full = reinterpret_cast<uint64_t*>(ds_smem);
    uint64_t* empty = full + DS_STAGES;
    uint64_t* rope_bar = empty + DS_STAGES;
    float* red = reinterpret_cast<float*>(ds_smem + DS_BAR_BYTES);
    float* fin = red + DS_CONSUMERS * 128;
    float* cos_s = fin + 128;
    float* sin_s = cos_s + DS_MAX_M * DS_ROPE_HALF;
    int* pg_s = reinterpret_cast<int*>(sin_s + DS_MAX_M * DS_ROPE_HALF);
    int* off_s = pg_s + DS_MAX_M;
    bf16* act = reinterpret_cast<bf16*>(off_s + DS_MAX_M);
    const int act_pitch = p.K + 8;
    constexpr uint32_t stage_bytes = DS_W_BYTES + (RESIDENT ? 0 : DS_A_BYTES);
    uint8_t* ring = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(act) + (RESIDENT ? (size_t)DS_MAX_M * act_pitch * 2 + (p.gamma != nullptr ? (size_t)p.K * 2 : 0) : 0) + 1023) &
        ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int s_begin = (int)((long long)p.total_stages * cta / G), s_end = (int)((long long)p.total_stages * (cta + 1) / G);
    const int n_my = s_end - s_begin;
    const int NS = p.nstages;

    if (threadIdx.x == 0) {
        for (int i = 0; i < DS_STAGES; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, DS_CONSUMERS);
        }
        mbar_init(rope_bar, 31);
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == DS_CONSUMERS) {
        // ------------------------------------------------------------------ producer (one thread)
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();   // every weight byte is used once per step
            const uint64_t keep = l2_policy_evict_last();
            auto issue_weights = [&](int it, int slot) {
                const int s = s_begin + it;
                const int tile = s / p.spt, ks = s % p.spt;
                // the full box always counts: rows past N and chunks past K are zero-filled by the TMA unit
                mbar_arrive_expect_tx(full + slot, stage_bytes);
                tma_load_3d_hint(ring + (size_t)slot * stage_bytes, &tmW, full + slot, 0, tile * DS_ROWS, ks * (DS_KW / 64), pol);
            };
            auto issue_acts = [&](int it, int slot) {
                const int s = s_begin + it;
                tma_load_3d_hint(ring + (size_t)slot * stage_bytes + DS_W_BYTES, &tmA, full + slot, 0, 0, (s % p.spt) * (DS_KW / 64), keep);
            };
            const int n_pre = min(NS, n_my);
            for (int it = 0; it < n_pre; ++it) issue_weights(it, it);   // static operands: before the dependency wait
            pdl_launch();
            pdl_wait();
            if (!RESIDENT)
                for (int it = 0; it < n_pre; ++it) issue_acts(it, it);
            int slot = 0, par = 0;   // ring position and wait parity, advanced incrementally (NS is a run-time value: no divisions)
            for (int it = n_pre; it < n_my; ++it) {
                mbar_wait(empty + slot, par);
                issue_weights(it, slot);
                if (!RESIDENT) issue_acts(it, slot);
                if (++slot == NS) { slot = 0; par ^= 1; }
            }
        } else {
            pdl_launch();
            if (p.epi == DS_EPI_ROPE_KV) {
                // lanes 1-31 (idle otherwise) stage what the epilogue needs per token -- cache page / offset and the cos / sin rows of
                // its position -- so that a tile's epilogue reads shared memory instead of three dependent global loads per element
                pdl_wait();
                const int half = p.hd >> 1;
                for (int i = lane - 1; i < p.M * half; i += 31) {
                    const int tk = i / half, j = i - tk * half;
                    const int pos = p.positions[tk];
                    cos_s[tk * DS_ROPE_HALF + j] = __bfloat162float(p.cos_t[(long long)pos * p.hd + j]);
                    sin_s[tk * DS_ROPE_HALF + j] = __bfloat162float(p.sin_t[(long long)pos * p.hd + j]);
                }
                for (int tk = lane - 1; tk < p.M; tk += 31) {
                    const int slot_i = p.slot_map[tk];
                    pg_s[tk] = slot_i / p.page;
                    off_s[tk] = slot_i % p.page;
                }
                mbar_arrive(rope_bar);
            }
        }
        // The HBM pipe would idle from here to the successor's first loads (this CTA's tail, the launch hand-over, the successor's
        // RMSNorm prologue): lanes 0-15 ask L2 for the rows of the stages the successor's CTA `cta` starts with.  Row segments of
        // a tile are contiguous ((kb - ka) KB per row), so a stage range costs 16 requests per tile touched.
        __syncwarp();
        if (p.pf_w != nullptr && cta < p.pf_grid && lane < DS_ROWS) {
            const int q0 = (int)((long long)p.pf_total * cta / p.pf_grid);
            const int q1 = min(q0 + p.pf_n, (int)((long long)p.pf_total * (cta + 1) / p.pf_grid));
            if (q1 > q0) {
                const int t_first = q0 / p.pf_spt, t_last = (q1 - 1) / p.pf_spt;
                for (int tl = t_first; tl <= t_last; ++tl) {
                    const int ka = (tl == t_first) ? q0 - tl * p.pf_spt : 0;
                    const int kb = (tl == t_last) ? q1 - tl * p.pf_spt : p.pf_spt;
                    const int row = tl * DS_ROWS + lane;
                    const int b0 = ka * (DS_KW * 2), b1 = min(kb * (DS_KW * 2), p.pf_rowbytes);
                    if (row < p.pf_N && b1 > b0) l2_prefetch_bulk(p.pf_w + (long long)row * p.pf_pitch + b0, (uint32_t)(b1 - b0));
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers (8 warps)
    pdl_launch();
    bf16* gamma_s = act + (size_t)DS_MAX_M * act_pitch;   // RMSNorm weight, staged before the dependency wait (static operand)
    if (RESIDENT && p.gamma != nullptr) {
        const uint4* g4 = reinterpret_cast<const uint4*>(p.gamma);
        for (int i = threadIdx.x; i < (p.K >> 3); i += DS_CONSUMERS * 32) reinterpret_cast<uint4*>(gamma_s)[i] = __ldg(g4 + i);
    }
    pdl_wait();
    const int g = lane >> 2, t = lane & 3;
    if (RESIDENT && p.gamma != nullptr) named_bar_sync(1, DS_CONSUMERS * 32);
    if (RESIDENT && (p.K >> 3) <= 32 * 20) {
        // one batch: the whole row of token m in this warp's registers (20 x 16 bytes per lane at K = 5120: a single L2 round trip),
        // sum of squares in the association of rmsnorm_kernel (lane-strided, in order), normalised straight from the registers
        const int nvec = p.K >> 3;
        for (int m = warp; m < DS_MAX_M; m += DS_CONSUMERS) {
            uint4* dst = reinterpret_cast<uint4*>(act + (size_t)m * act_pitch);
            if (m >= p.M) {
                for (int i = lane; i < nvec; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
                continue;
            }
            const uint4* xr = reinterpret_cast<const uint4*>(p.a + (long long)m * p.lda);
            uint4 q[20];
#pragma unroll
            for (int u = 0; u < 20; ++u) q[u] = (lane + 32 * u < nvec) ? __ldcg(xr + lane + 32 * u) : make_uint4(0, 0, 0, 0);
            if (p.gamma == nullptr) {
#pragma unroll
                for (int u = 0; u < 20; ++u)
                    if (lane + 32 * u < nvec) dst[lane + 32 * u] = q[u];
                continue;
            }
            float ss = 0.f;
#pragma unroll
            for (int u = 0; u < 20; ++u) {
                const float2 a = unpack_bf16x2(q[u].x), b = unpack_bf16x2(q[u].y), c = unpack_bf16x2(q[u].z), d = unpack_bf16x2(q[u].w);
                ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
            }
            const float rstd = rsqrtf(warp_sum(ss) / (float)p.K + p.eps);
#pragma unroll
            for (int u = 0; u < 20; ++u) {
                const int i = lane + 32 * u;
                if (i < nvec) {
                    const uint4 gm = reinterpret_cast<const uint4*>(gamma_s)[i];
                    const uint32_t xi[4] = {q[u].x, q[u].y, q[u].z, q[u].w}, gi[4] = {gm.x, gm.y, gm.z, gm.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]);
                        o[j] = pack_bf16x2(gv.x * bf16_round(xv.x * rstd), gv.y * bf16_round(xv.y * rstd));
                    }
                    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        named_bar_sync(1, DS_CONSUMERS * 32);
    } else if (RESIDENT) {
        // warp m stages token m: raw copy + sum of squares (the association of rmsnorm_kernel: lane-strided, in order), then
        // normalise in place.  Loads go out in batches of 10 per lane so that a 5120-wide row costs two L2 round trips.
        const int nvec = p.K >> 3;
        for (int m = warp; m < DS_MAX_M; m += DS_CONSUMERS) {
            uint4* dst = reinterpret_cast<uint4*>(act + (size_t)m * act_pitch);
            if (m >= p.M) {
                for (int i = lane; i < nvec; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
                continue;
            }
            const uint4* xr = reinterpret_cast<const uint4*>(p.a + (long long)m * p.lda);
            float ss = 0.f;
            for (int i0 = lane; i0 < nvec; i0 += 320) {
                uint4 q[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) q[u] = (i0 + 32 * u < nvec) ? __ldcg(xr + i0 + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    if (i0 + 32 * u < nvec) dst[i0 + 32 * u] = q[u];
                    const float2 a = unpack_bf16x2(q[u].x), b = unpack_bf16x2(q[u].y), c = unpack_bf16x2(q[u].z), d = unpack_bf16x2(q[u].w);
                    ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
                }
            }
            if (p.gamma != nullptr) {
                const float rstd = rsqrtf(warp_sum(ss) / (float)p.K + p.eps);
                const uint4* g4 = reinterpret_cast<const uint4*>(gamma_s);
                for (int i = lane; i < nvec; i += 32) {
                    const uint4 q = dst[i], gm = g4[i];
                    const uint32_t xi[4] = {q.x, q.y, q.z, q.w}, gi[4] = {gm.x, gm.y, gm.z, gm.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]);
                        o[j] = pack_bf16x2(gv.x * bf16_round(xv.x * rstd), gv.y * bf16_round(xv.y * rstd));
                    }
                    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        named_bar_sync(1, DS_CONSUMERS * 32);
    }

    float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};   // two accumulators: independent HMMA chains
    // ldmatrix lane addressing.  A (weights): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15).
    // A stage holds 8 slabs (one per 64-wide k-chunk) of [16 rows x 128 B] with the 128-byte swizzle: 16-byte unit u of row r
    // sits at r * 128 + ((u ^ (r & 7)) << 4).  Warp w takes the k16-steps w, w + 8, w + 16, w + 24 of a stage: slab
    // (w >> 2) + 2 i, units 2 (w & 3) and 2 (w & 3) + 1 -- the in-slab offsets are per-lane constants.
    const int a_row = ((lane >> 3) & 1) * 8 + (lane & 7), a_half = lane >> 4;
    // B (tokens, [token][k]): matrices (tokens 0-7, k 0-7), (tokens 0-7, k 8-15); lanes 16-31 repeat valid addresses
    const int b_row = lane & 7, b_half = (lane >> 3) & 1;
    const int u_w = (warp & 3) * 2;
    const uint32_t a_off = a_row * 128 + (((u_w + a_half) ^ (a_row & 7)) << 4) + (warp >> 2) * 2048;
    const uint32_t b_off_str = DS_W_BYTES + b_row * 128 + (((u_w + b_half) ^ b_row) << 4) + (warp >> 2) * 1024;
    const bf16* b_res0 = act + (size_t)b_row * act_pitch + b_half * 8 + warp * 16;

    int tile = s_begin / p.spt, ks = s_begin - tile * p.spt;
    int slot = -1, par = 1;
    for (int it = 0; it < n_my; ++it) {
        if (++slot == NS) slot = 0;
        if (slot == 0) par ^= 1;
        const int k0 = ks * DS_KW;
        const int kw = min(DS_KW, p.K - k0);
        mbar_wait(full + slot, par);
        const uint8_t* st = ring + (size_t)slot * stage_bytes;
        if (kw == DS_KW) {
            uint32_t af[4][4], bq[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], st + a_off + i * 4096);
                if (RESIDENT) ldmatrix_x2(bq[i][0], bq[i][1], b_res0 + k0 + i * 128);
                else ldmatrix_x2(bq[i][0], bq[i][1], st + b_off_str + i * 2048);
            }
            mma_bf16_16816(c, af[0], bq[0][0], bq[0][1]);
            mma_bf16_16816(c2, af[1], bq[1][0], bq[1][1]);
            mma_bf16_16816(c, af[2], bq[2][0], bq[2][1]);
            mma_bf16_16816(c2, af[3], bq[3][0], bq[3][1]);
        } else {
            const int n16 = kw >> 4;
            for (int j = warp; j < n16; j += DS_CONSUMERS) {
                uint32_t af[4], b0, b1;
                const int i = j >> 3;
                ldmatrix_x4(af[0], af[1], af[2], af[3], st + a_off + i * 4096);
                if (RESIDENT) ldmatrix_x2(b0, b1, b_res0 + k0 + i * 128);
                else ldmatrix_x2(b0, b1, st + b_off_str + i * 2048);
                mma_bf16_16816(c, af, b0, b1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);

        const bool tile_end = (ks == p.spt - 1), range_end = (it == n_my - 1);
        const int tile_now = tile;
        if (++ks == p.spt) { ks = 0; ++tile; }
        if (!tile_end && !range_end) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i] += c2[i]; c2[i] = 0.f; }
        // ---- fold the eight partial tiles (warp order), then finish or hand over the tile
        float* pw = red + warp * 128;
        pw[g * 8 + 2 * t] = c[0];
        pw[g * 8 + 2 * t + 1] = c[1];
        pw[(g + 8) * 8 + 2 * t] = c[2];
        pw[(g + 8) * 8 + 2 * t + 1] = c[3];
        c[0] = c[1] = c[2] = c[3] = 0.f;
        named_bar_sync(1, DS_CONSUMERS * 32);
        const int e = threadIdx.x;                       // 0..255; the first 128 threads own one (row, token) each
        const int rl = e & 15, tok = (e >> 4) & 7;
        const bool head_part = (tile_now * p.spt < s_begin);          // this CTA does not hold the tile's first stage: contributor
        const bool cut_tail = (!tile_end);                        // the tile continues in the next CTA: this CTA finishes it
        float x = 0.f;
        if (e < 128) {
#pragma unroll
            for (int w = 0; w < DS_CONSUMERS; ++w) x += red[w * 128 + rl * 8 + tok];
        }
        if (head_part) {
            if (e < 128) p.partial[(size_t)cta * 128 + e] = x;
            named_bar_sync(1, DS_CONSUMERS * 32);
            if (e == 0) {   // the barrier made the 128 stores visible to this thread; its fence + release publishes them
                __threadfence();
                st_release(p.flags + cta, 1);
            }
            continue;
        }
        if (cut_tail) {
            if (e == 0) {
                const uint64_t t0 = global_timer_ns();
                while (ld_acquire(p.flags + cta + 1) == 0) {
                    if (global_timer_ns() - t0 > 5000000000ull) {
                        printf("ivlm: decode_stream partial-tile wait timeout cta=%d\n", cta);
                        __trap();
                    }
                }
            }
            named_bar_sync(1, DS_CONSUMERS * 32);
            if (e < 128) x += __ldcg(p.partial + (size_t)(cta + 1) * 128 + e);
            named_bar_sync(1, DS_CONSUMERS * 32);
            if (e == 0) p.flags[cta + 1] = 0;   // self-resetting: the next launch finds it clear (stream order)
        }
        // ---- epilogue
        const int row0 = tile_now * DS_ROWS;
        if (p.epi == DS_EPI_PLAIN) {
            const int row = row0 + rl;
            if (e < 128 && row < p.N && tok < p.M) {
                if (p.bias != nullptr) x += __bfloat162float(p.bias[row]);
                if (p.round_steps) x = bf16_round(x);
                if (p.act != ACT_NONE) {
                    x = apply_act(x, p.act);
                    if (p.round_steps) x = bf16_round(x);
                }
                if (p.res != nullptr) x += __bfloat162float(p.res[(long long)tok * p.ldr + row]);
                const long long oi = (long long)tok * p.ldo + row;
                if (p.out_f32) reinterpret_cast<float*>(p.out)[oi] = x;
                else reinterpret_cast<bf16*>(p.out)[oi] = __float2bfloat16_rn(x);
            }
        } else {
            if (e < 128) fin[rl * 8 + tok] = bf16_round(x);   // the Linear output as the reference holds it (bf16)
            named_bar_sync(1, DS_CONSUMERS * 32);
            if (e < 64 && (e >> 3) < p.M) {
                const int j8 = e & 7, tk = e >> 3;            // 8 features x 8 tokens
                const float lo = fin[j8 * 8 + tk], hi = fin[(j8 + 8) * 8 + tk];
                if (p.epi == DS_EPI_SWIGLU) {
                    // rows 0-7: gate of features 8*tile + j8, rows 8-15: up of the same features
                    const float v = bf16_round(apply_act(lo, ACT_SILU)) * hi;
                    reinterpret_cast<bf16*>(p.out)[(long long)tk * p.ldo + tile_now * 8 + j8] = __float2bfloat16_rn(v);
                } else {
                    const int D = p.H * p.hd, half = p.hd >> 1;
                    const int sec = row0 / D, r = row0 - sec * D;
                    mbar_wait(rope_bar, 0);   // completes once per launch; later waits return at once
                    const long long pg = pg_s[tk], off = off_s[tk];
                    if (sec < 2) {
                        const int hh = r / p.hd, j = ((r % p.hd) >> 4) * 8 + j8;   // rotary pair (j, j + half) of head hh
                        const float cs = cos_s[tk * DS_ROPE_HALF + j];
                        const float sn = sin_s[tk * DS_ROPE_HALF + j];
                        const bf16 o1 = __float2bfloat16_rn(bf16_round(lo * cs) + bf16_round(-hi * sn));
                        const bf16 o2 = __float2bfloat16_rn(bf16_round(hi * cs) + bf16_round(lo * sn));
                        if (sec == 0) {
                            bf16* q = reinterpret_cast<bf16*>(p.out) + (long long)tk * p.ldo + hh * p.hd;
                            q[j] = o1;
                            q[j + half] = o2;
                        } else {
                            bf16* kc = p.k_cache + ((pg * p.H + hh) * p.page + off) * p.hd;
                            kc[j] = o1;
                            kc[j + half] = o2;
                        }
                    } else {
                        // v rows are in natural order: 16 consecutive features of one head
                        const int f0 = r + j8;
                        bf16* vc0 = p.v_cache + ((pg * p.H + f0 / p.hd) * p.page + off) * p.hd;
                        vc0[f0 % p.hd] = __float2bfloat16_rn(lo);
                        const int f1 = f0 + 8;
                        bf16* vc1 = p.v_cache + ((pg * p.H + f1 / p.hd) * p.page + off) * p.hd;
                        vc1[f1 % p.hd] = __float2bfloat16_rn(hi);
                    }
                }
            }
        }
        named_bar_sync(1, DS_CONSUMERS * 32);   // red / fin are free again
    }
}


// ------------------------------------------------------------------------------------------------ chained launches
// Several weight-streaming layers that depend on each other (o_proj -> gate/up -> down_proj -> the next layer's qkv) as ONE
// launch: the phases are separated by grid-wide barriers instead of kernel boundaries, and the producer thread simply keeps
// going -- while the consumers of every CTA fold their last tiles of phase i, run its epilogue, meet at the barrier and run the
// RMSNorm prologue of phase i+1, the ring already holds (and HBM keeps delivering) the first stages of phase i+1's weights.
// Separate launches pay ~6 us each for tail + launch hand-over + ramp-up against 8-43 us of streaming.
// Every phase keeps the tile / stage decomposition, fold order and epilogue of decode_stream_kernel: bit-identical results.
//   * ring: uniform 24 KB stages (16 KB weights + 8 KB of streamed activation, unused by resident phases), depth chosen to fit
//   * a streamed phase's activation is another phase's output: its weight loads are issued ahead, the activation loads of the
//     same stages follow once the barrier has been passed (the stage's full barrier expects both);
//   * barrier g counts the CTAs that finished phase g (release: bar.sync + fence + atomic); the last arriver at barrier g
//     re-arms barrier g-1 (every CTA has passed it by then), the last one re-arms itself: no host-side reset between launches.
// All CTAs must be resident at once (grid <= SMs, one CTA per SM by shared memory); spins trap after 2 s instead of hanging.
constexpr int DC_MAX_PH = 4;
struct ChainMaps { CUtensorMap w[DC_MAX_PH]; CUtensorMap a[DC_MAX_PH]; };
struct DecodeChainParams {
    DecodeStreamParams ph[DC_MAX_PH];
    int n;            // phases
    int nstages;      // ring depth
    int max_k_res;    // largest K among the resident phases (layout of the activation / RMSNorm-weight area)
    int* gbar;        // [DC_MAX_PH] phase barriers, zero between launches
};

IVLM_DEVINL void gbar_spin(const int* bar, int target) {
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire(bar) < target) {
        if (global_timer_ns() - t0 > 2000000000ull) {
            printf("ivlm: decode_chain grid barrier timeout cta=%d (have %d of %d)\n", (int)blockIdx.x, ld_acquire(bar), target);
            __trap();
        }
    }
}

__global__ void __launch_bounds__(DS_THREADS, 1)
decode_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ DecodeChainParams cp) {
    extern __shared__ __align__(128) uint8_t ds_smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(ds_smem);
    uint64_t* empty = full + DS_STAGES;
    uint64_t* rope_bar = empty + DS_STAGES;
    float* red = reinterpret_cast<float*>(ds_smem + DS_BAR_BYTES);
    float* fin = red + DS_CONSUMERS * 128;
    float* cos_s = fin + 128;
    float* sin_s = cos_s + DS_MAX_M * DS_ROPE_HALF;
    int* pg_s = reinterpret_cast<int*>(sin_s + DS_MAX_M * DS_ROPE_HALF);
    int* off_s = pg_s + DS_MAX_M;
    bf16* act = reinterpret_cast<bf16*>(off_s + DS_MAX_M);
    bf16* gamma_s = act + (size_t)DS_MAX_M * (cp.max_k_res + 8);
    constexpr uint32_t stage_bytes = DS_W_BYTES + DS_A_BYTES;
    uint8_t* ring = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(gamma_s) + (size_t)cp.max_k_res * 2 + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int NS = cp.nstages;

    if (threadIdx.x == 0) {
        for (int i = 0; i < DS_STAGES; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, DS_CONSUMERS);
        }
        mbar_init(rope_bar, 31);
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == DS_CONSUMERS) {
        if (lane == 0) {
            // ------------------------------------------------------------------ producer (one thread), all phases back to back
            const uint64_t pol = l2_policy_evict_first();
            const uint64_t keep = l2_policy_evict_last();
            int slot = 0, par = 0;
            long long issued = 0;          // stages issued so far over all phases
            bool waited = false;
            pdl_launch();
            for (int ph = 0; ph < cp.n; ++ph) {
                const DecodeStreamParams& p = cp.ph[ph];
                const int Gp = G < p.tiles ? G : p.tiles;   // CTAs taking part in this phase (never more than whole tiles)
                if (cta >= Gp) continue;
                const int s_begin = (int)((long long)p.total_stages * cta / Gp), s_end = (int)((long long)p.total_stages * (cta + 1) / Gp);
                const bool streamed = !p.resident;
                bool a_ready = !streamed;
                int pend_it[DS_STAGES], pend_slot[DS_STAGES], npend = 0;
                auto acts_ready = [&]() {
                    // the activation of this phase: the launch's input (phase 0: wait for the predecessor kernel) or the output of
                    // the previous phase (grid barrier ph-1), written with generic stores and read here through the async proxy
                    if (ph == 0) { if (!waited) { pdl_wait(); waited = true; } }
                    else gbar_spin(cp.gbar + ph - 1, G);
                    asm volatile("fence.proxy.async;" ::: "memory");
                    for (int i = 0; i < npend; ++i)
                        tma_load_3d_hint(ring + (size_t)pend_slot[i] * stage_bytes + DS_W_BYTES, &maps.a[ph], full + pend_slot[i], 0, 0,
                                         ((s_begin + pend_it[i]) % p.spt) * (DS_KW / 64), keep);
                    npend = 0;
                    a_ready = true;
                };
                for (int it = 0; it < s_end - s_begin; ++it) {
                    if (issued >= NS) {
                        if (!a_ready && npend == NS) acts_ready();   // every slot holds a stage of this phase that lacks its activation
                        mbar_wait(empty + slot, par ^ 1);
                    }
                    const int s = s_begin + it;
                    const int tile = s / p.spt, ks = s % p.spt;
                    mbar_arrive_expect_tx(full + slot, streamed ? stage_bytes : (uint32_t)DS_W_BYTES);
                    tma_load_3d_hint(ring + (size_t)slot * stage_bytes, &maps.w[ph], full + slot, 0, tile * DS_ROWS, ks * (DS_KW / 64), pol);
                    if (streamed) {
                        if (a_ready)
                            tma_load_3d_hint(ring + (size_t)slot * stage_bytes + DS_W_BYTES, &maps.a[ph], full + slot, 0, 0, ks * (DS_KW / 64), keep);
                        else { pend_it[npend] = it; pend_slot[npend] = slot; ++npend; }
                    }
                    ++issued;
                    if (++slot == NS) { slot = 0; par ^= 1; }
                }
                if (!a_ready) acts_ready();
            }
            if (!waited) pdl_wait();
        } else {
            pdl_launch();
            // lanes 1-31: rotary tables / cache slots of the (single) ROPE_KV phase -- inputs of the launch, not of a phase
            for (int ph = 0; ph < cp.n; ++ph) {
                const DecodeStreamParams& p = cp.ph[ph];
                if (p.epi != DS_EPI_ROPE_KV) continue;
                pdl_wait();
                const int half = p.hd >> 1;
                for (int i = lane - 1; i < p.M * half; i += 31) {
                    const int tk = i / half, j = i - tk * half;
                    const int pos = p.positions[tk];
                    cos_s[tk * DS_ROPE_HALF + j] = __bfloat162float(p.cos_t[(long long)pos * p.hd + j]);
                    sin_s[tk * DS_ROPE_HALF + j] = __bfloat162float(p.sin_t[(long long)pos * p.hd + j]);
                }
                for (int tk = lane - 1; tk < p.M; tk += 31) {
                    const int slot_i = p.slot_map[tk];
                    pg_s[tk] = slot_i / p.page;
                    off_s[tk] = slot_i % p.page;
                }
                mbar_arrive(rope_bar);
                break;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers (8 warps)
    pdl_launch();
    const int g = lane >> 2, t = lane & 3;
    const int a_row = ((lane >> 3) & 1) * 8 + (lane & 7), a_half = lane >> 4;
    const int b_row = lane & 7, b_half = (lane >> 3) & 1;
    const int u_w = (warp & 3) * 2;
    const uint32_t a_off = a_row * 128 + (((u_w + a_half) ^ (a_row & 7)) << 4) + (warp >> 2) * 2048;
    const uint32_t b_off_str = DS_W_BYTES + b_row * 128 + (((u_w + b_half) ^ b_row) << 4) + (warp >> 2) * 1024;
    int slot = -1, par = 1;
    for (int ph = 0; ph < cp.n; ++ph) {
        const DecodeStreamParams& p = cp.ph[ph];
        const bool resident = p.resident != 0;
        const int act_pitch = p.K + 8;
        // ---- static operand of the prologue first, then the dependency: the predecessor kernel / the previous phase
        if (resident && p.gamma != nullptr) {
            const uint4* g4 = reinterpret_cast<const uint4*>(p.gamma);
            for (int i = threadIdx.x; i < (p.K >> 3); i += DS_CONSUMERS * 32) reinterpret_cast<uint4*>(gamma_s)[i] = __ldg(g4 + i);
        }
        if (ph == 0) pdl_wait();
        else {
            if (threadIdx.x == 0) gbar_spin(cp.gbar + ph - 1, G);
        }
        named_bar_sync(1, DS_CONSUMERS * 32);
        if (resident) {
            const int nvec = p.K >> 3;
            for (int m = warp; m < DS_MAX_M; m += DS_CONSUMERS) {
                uint4* dst = reinterpret_cast<uint4*>(act + (size_t)m * act_pitch);
                if (m >= p.M) {
                    for (int i = lane; i < nvec; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
                    continue;
                }
                const uint4* xr = reinterpret_cast<const uint4*>(p.a + (long long)m * p.lda);
                float ss = 0.f;
                for (int i0 = lane; i0 < nvec; i0 += 640) {   // sum of squares in the association of rmsnorm_kernel
                    uint4 q[20];
#pragma unroll
                    for (int u = 0; u < 20; ++u) q[u] = (i0 + 32 * u < nvec) ? __ldcg(xr + i0 + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 20; ++u) {
                        if (i0 + 32 * u < nvec) dst[i0 + 32 * u] = q[u];
                        const float2 a = unpack_bf16x2(q[u].x), b = unpack_bf16x2(q[u].y), c = unpack_bf16x2(q[u].z), d = unpack_bf16x2(q[u].w);
                        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
                    }
                }
                if (p.gamma != nullptr) {
                    const float rstd = rsqrtf(warp_sum(ss) / (float)p.K + p.eps);
                    const uint4* g4 = reinterpret_cast<const uint4*>(gamma_s);
                    for (int i = lane; i < nvec; i += 32) {
                        const uint4 q = dst[i], gm = g4[i];
                        const uint32_t xi[4] = {q.x, q.y, q.z, q.w}, gi[4] = {gm.x, gm.y, gm.z, gm.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]);
                            o[j] = pack_bf16x2(gv.x * bf16_round(xv.x * rstd), gv.y * bf16_round(xv.y * rstd));
                        }
                        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            named_bar_sync(1, DS_CONSUMERS * 32);
        }
        const int Gp = G < p.tiles ? G : p.tiles;
        const int s_begin = cta < Gp ? (int)((long long)p.total_stages * cta / Gp) : 0;
        const int n_my = cta < Gp ? (int)((long long)p.total_stages * (cta + 1) / Gp) - s_begin : 0;
        const bf16* b_res0 = act + (size_t)b_row * act_pitch + b_half * 8 + warp * 16;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
        int tile = s_begin / p.spt, ks = s_begin - tile * p.spt;
        for (int it = 0; it < n_my; ++it) {
            if (++slot == NS) slot = 0;
            if (slot == 0) par ^= 1;
            const int k0 = ks * DS_KW;
            const int kw = min(DS_KW, p.K - k0);
            mbar_wait(full + slot, par);
            const uint8_t* st = ring + (size_t)slot * stage_bytes;
            if (kw == DS_KW) {
                uint32_t af[4][4], bq[4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], st + a_off + i * 4096);
                    if (resident) ldmatrix_x2(bq[i][0], bq[i][1], b_res0 + k0 + i * 128);
                    else ldmatrix_x2(bq[i][0], bq[i][1], st + b_off_str + i * 2048);
                }
                mma_bf16_16816(c, af[0], bq[0][0], bq[0][1]);
                mma_bf16_16816(c2, af[1], bq[1][0], bq[1][1]);
                mma_bf16_16816(c, af[2], bq[2][0], bq[2][1]);
                mma_bf16_16816(c2, af[3], bq[3][0], bq[3][1]);
            } else {
                const int n16 = kw >> 4;
                for (int j = warp; j < n16; j += DS_CONSUMERS) {
                    uint32_t af[4], b0, b1;
                    const int i = j >> 3;
                    ldmatrix_x4(af[0], af[1], af[2], af[3], st + a_off + i * 4096);
                    if (resident) ldmatrix_x2(b0, b1, b_res0 + k0 + i * 128);
                    else ldmatrix_x2(b0, b1, st + b_off_str + i * 2048);
                    mma_bf16_16816(c, af, b0, b1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);

            const bool tile_end = (ks == p.spt - 1), range_end = (it == n_my - 1);
            const int tile_now = tile;
            if (++ks == p.spt) { ks = 0; ++tile; }
            if (!tile_end && !range_end) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) { c[i] += c2[i]; c2[i] = 0.f; }
            float* pw = red + warp * 128;
            pw[g * 8 + 2 * t] = c[0];
            pw[g * 8 + 2 * t + 1] = c[1];
            pw[(g + 8) * 8 + 2 * t] = c[2];
            pw[(g + 8) * 8 + 2 * t + 1] = c[3];
            c[0] = c[1] = c[2] = c[3] = 0.f;
            named_bar_sync(1, DS_CONSUMERS * 32);
            const int e = threadIdx.x;
            const int rl = e & 15, tok = (e >> 4) & 7;
            const bool head_part = (tile_now * p.spt < s_begin);
            const bool cut_tail = (!tile_end);
            float x = 0.f;
            if (e < 128) {
#pragma unroll
                for (int w = 0; w < DS_CONSUMERS; ++w) x += red[w * 128 + rl * 8 + tok];
            }
            if (head_part) {
                if (e < 128) p.partial[(size_t)cta * 128 + e] = x;
                named_bar_sync(1, DS_CONSUMERS * 32);
                if (e == 0) {
                    __threadfence();
                    st_release(p.flags + cta, 1);
                }
                continue;
            }
            if (cut_tail) {
                if (e == 0) {
                    const uint64_t t0 = global_timer_ns();
                    while (ld_acquire(p.flags + cta + 1) == 0) {
                        if (global_timer_ns() - t0 > 2000000000ull) {
                            printf("ivlm: decode_chain partial-tile wait timeout cta=%d phase=%d\n", cta, ph);
                            __trap();
                        }
                    }
                }
                named_bar_sync(1, DS_CONSUMERS * 32);
                if (e < 128) x += __ldcg(p.partial + (size_t)(cta + 1) * 128 + e);
                named_bar_sync(1, DS_CONSUMERS * 32);
                if (e == 0) p.flags[cta + 1] = 0;
            }
            const int row0 = tile_now * DS_ROWS;
            if (p.epi == DS_EPI_PLAIN) {
                const int row = row0 + rl;
                if (e < 128 && row < p.N && tok < p.M) {
                    if (p.bias != nullptr) x += __bfloat162float(p.bias[row]);
                    if (p.round_steps) x = bf16_round(x);
                    if (p.act != ACT_NONE) {
                        x = apply_act(x, p.act);
                        if (p.round_steps) x = bf16_round(x);
                    }
                    if (p.res != nullptr) x += __bfloat162float(__ldcg(p.res + (long long)tok * p.ldr + row));
                    const long long oi = (long long)tok * p.ldo + row;
                    if (p.out_f32) reinterpret_cast<float*>(p.out)[oi] = x;
                    else reinterpret_cast<bf16*>(p.out)[oi] = __float2bfloat16_rn(x);
                }
            } else {
                if (e < 128) fin[rl * 8 + tok] = bf16_round(x);
                named_bar_sync(1, DS_CONSUMERS * 32);
                if (e < 64 && (e >> 3) < p.M) {
                    const int j8 = e & 7, tk = e >> 3;
                    const float lo = fin[j8 * 8 + tk], hi = fin[(j8 + 8) * 8 + tk];
                    if (p.epi == DS_EPI_SWIGLU) {
                        const float v = bf16_round(apply_act(lo, ACT_SILU)) * hi;
                        reinterpret_cast<bf16*>(p.out)[(long long)tk * p.ldo + tile_now * 8 + j8] = __float2bfloat16_rn(v);
                    } else {
                        const int D = p.H * p.hd, half = p.hd >> 1;
                        const int sec = row0 / D, r = row0 - sec * D;
                        mbar_wait(rope_bar, 0);
                        const long long pg = pg_s[tk], off = off_s[tk];
                        if (sec < 2) {
                            const int hh = r / p.hd, j = ((r % p.hd) >> 4) * 8 + j8;
                            const float cs = cos_s[tk * DS_ROPE_HALF + j];
                            const float sn = sin_s[tk * DS_ROPE_HALF + j];
                            const bf16 o1 = __float2bfloat16_rn(bf16_round(lo * cs) + bf16_round(-hi * sn));
                            const bf16 o2 = __float2bfloat16_rn(bf16_round(hi * cs) + bf16_round(lo * sn));
                            if (sec == 0) {
                                bf16* q = reinterpret_cast<bf16*>(p.out) + (long long)tk * p.ldo + hh * p.hd;
                                q[j] = o1;
                                q[j + half] = o2;
                            } else {
                                bf16* kc = p.k_cache + ((pg * p.H + hh) * p.page + off) * p.hd;
                                kc[j] = o1;
                                kc[j + half] = o2;
                            }
                        } else {
                            const int f0 = r + j8;
                            bf16* vc0 = p.v_cache + ((pg * p.H + f0 / p.hd) * p.page + off) * p.hd;
                            vc0[f0 % p.hd] = __float2bfloat16_rn(lo);
                            const int f1 = f0 + 8;
                            bf16* vc1 = p.v_cache + ((pg * p.H + f1 / p.hd) * p.page + off) * p.hd;
                            vc1[f1 % p.hd] = __float2bfloat16_rn(hi);
                        }
                    }
                }
            }
            named_bar_sync(1, DS_CONSUMERS * 32);
        }
        // ---- this CTA's part of phase ph is in global memory: arrive at barrier ph (the last phase's barrier is only counted)
        named_bar_sync(1, DS_CONSUMERS * 32);
        if (threadIdx.x == 0) {
            __threadfence();
            const int old = atomicAdd(cp.gbar + ph, 1);
            if (old == G - 1) {
                if (ph > 0) cp.gbar[ph - 1] = 0;        // everybody has passed barrier ph-1
                if (ph == cp.n - 1) cp.gbar[ph] = 0;    // nobody waits on the last one
            }
        }
    }
}

}  // namespace ivlm

using namespace ivlm;

// argument checks and the fields shared by the single launch and the chained one
static int fill_decode_params(ivlm_handle h, const ivlm_decode_linear_args* a, DecodeStreamParams& p) {
    IVLM_REQUIRE(h && a && a->a && a->w && a->out, "decode_linear: null argument");
    IVLM_REQUIRE(a->M >= 1 && a->M <= DS_MAX_M, "decode_linear: token count %d outside 1..%d", a->M, DS_MAX_M);
    IVLM_REQUIRE(a->N > 0 && a->K >= 64 && a->K % 64 == 0, "decode_linear: N=%d K=%d (K must be a multiple of 64)", a->N, a->K);
    IVLM_REQUIRE((a->lda * 2) % 16 == 0 && (a->ldw * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->a) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
                 "decode_linear: operands must be 16-byte aligned with 16-byte row pitches");
    IVLM_REQUIRE(a->epilogue >= DS_EPI_PLAIN && a->epilogue <= DS_EPI_ROPE_KV, "decode_linear: unknown epilogue %d", a->epilogue);
    IVLM_REQUIRE(h->ws != nullptr && h->ws_bytes >= IVLM_WS_COUNTER_BYTES + (size_t)(h->num_sms + 1) * 512,
                 "decode_linear: bind a workspace first (ivlm_set_workspace)");
    p = DecodeStreamParams{};
    p.a = reinterpret_cast<const bf16*>(a->a); p.lda = a->lda;
    p.w = reinterpret_cast<const bf16*>(a->w); p.ldw = a->ldw;
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.gamma = reinterpret_cast<const bf16*>(a->norm_gamma); p.eps = a->norm_eps;
    p.epi = a->epilogue; p.act = a->act;
    p.out_f32 = a->out_dtype == IVLM_F32;
    p.round_steps = a->out_dtype == IVLM_BF16 ? 1 : 0;
    p.bias = reinterpret_cast<const bf16*>(a->bias);
    p.res = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
    p.out = a->out; p.ldo = a->ldo;
    if (a->epilogue == DS_EPI_SWIGLU) {
        IVLM_REQUIRE(a->N % 16 == 0 && !a->bias && !a->residual && a->out_dtype == IVLM_BF16,
                     "decode_linear: SWIGLU needs interleaved gate/up rows (N %% 16 == 0), bf16 output, no bias/residual");
    }
    if (a->epilogue == DS_EPI_ROPE_KV) {
        IVLM_REQUIRE(a->positions && a->slot_map && a->cos_t && a->sin_t && a->k_cache && a->v_cache && a->H > 0 && a->hd > 0 &&
                         a->hd % 16 == 0 && a->hd <= 2 * DS_ROPE_HALF && a->N == 3 * a->H * a->hd && a->page_size > 0 && a->out_dtype == IVLM_BF16 && !a->bias &&
                         !a->residual,
                     "decode_linear: ROPE_KV needs N == 3*H*hd, hd %% 16 == 0, hd <= 128, tables, cache and slot map, bf16 q output");
        p.positions = a->positions; p.slot_map = a->slot_map;
        p.cos_t = reinterpret_cast<const bf16*>(a->cos_t); p.sin_t = reinterpret_cast<const bf16*>(a->sin_t);
        p.k_cache = reinterpret_cast<bf16*>(a->k_cache); p.v_cache = reinterpret_cast<bf16*>(a->v_cache);
        p.H = a->H; p.hd = a->hd; p.page = a->page_size;
    }
    p.tiles = (a->N + DS_ROWS - 1) / DS_ROWS;
    p.spt = (a->K + DS_KW - 1) / DS_KW;
    p.total_stages = p.tiles * p.spt;
    p.partial = reinterpret_cast<float*>(h->ws + IVLM_WS_COUNTER_BYTES);
    p.flags = reinterpret_cast<int*>(h->ws + DS_FLAG_OFFSET_BYTES);
    return IVLM_OK;
}

extern "C" int ivlm_decode_linear(ivlm_handle h, const ivlm_decode_linear_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DecodeStreamParams p;
    IVLM_TRY(fill_decode_params(h, a, p));
    // shared memory: barriers + fold buffers, the resident activation (when it fits next to the ring), the 1024-byte aligned ring
    const size_t fixed = DS_FIXED_BYTES;
    const size_t act_bytes = (size_t)DS_MAX_M * (a->K + 8) * 2 + (a->norm_gamma != nullptr ? (size_t)a->K * 2 : 0);   // + staged RMSNorm weight
    // ring depth: 7 stages when they fit next to the resident activation (133.1-133.4 us per 13B layer against 134.4-134.9 with 6 in
    // three runs of tools/prof_decode.py; 8 no longer fits with the staged RMSNorm weight), option "ds_stages" overrides
    int ns = (h->ds_stages >= 2 && h->ds_stages <= DS_STAGES) ? h->ds_stages : 7;
    {
        const size_t cap_ = 227 * 1024;
        const size_t act_ = (size_t)DS_MAX_M * (a->K + 8) * 2 + (a->norm_gamma != nullptr ? (size_t)a->K * 2 : 0);
        const bool res_ = !(h->ds_force_stream && a->norm_gamma == nullptr) && DS_FIXED_BYTES + act_ + 1024 + (size_t)6 * DS_W_BYTES <= cap_;
        if (h->ds_stages == 0) {
            const size_t need7 = DS_FIXED_BYTES + 1024 + (res_ ? act_ + (size_t)7 * DS_W_BYTES : (size_t)7 * (DS_W_BYTES + DS_A_BYTES));
            if (need7 > cap_) ns = 6;
        }
    }
    p.nstages = ns;
    const size_t ring_res = (size_t)ns * DS_W_BYTES, ring_str = (size_t)ns * (DS_W_BYTES + DS_A_BYTES);
    const size_t cap = 227 * 1024;
    p.resident = (fixed + act_bytes + 1024 + ring_res <= cap) ? 1 : 0;
    if (h->ds_force_stream && a->norm_gamma == nullptr) p.resident = 0;
    IVLM_REQUIRE(p.resident || a->norm_gamma == nullptr, "decode_linear: fused RMSNorm needs K <= %d (activation resident in shared memory)",
                 (int)((cap - fixed - 1024 - ring_res) / (2 * (DS_MAX_M + 1)) - 8));
    const size_t smem = fixed + 1024 + (p.resident ? act_bytes + ring_res : ring_str);
    const CUtensorMap *tw, *ta;
    IVLM_TRY(get_tmap_bf16_kchunk3d(h, a->w, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, DS_ROWS, &tw));
    IVLM_TRY(get_tmap_bf16_kchunk3d(h, a->a, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, DS_MAX_M, &ta));
    // one CTA per SM, but never more CTAs than whole tiles: a range then always reaches the end of the tile it starts in
    int grid = h->num_sms < p.tiles ? h->num_sms : p.tiles;
    if (a->prefetch_w != nullptr && a->prefetch_N > 0 && a->prefetch_K > 0 && h->ds_prefetch_kb != 0) {
        IVLM_REQUIRE((reinterpret_cast<uintptr_t>(a->prefetch_w) & 15) == 0 && (a->prefetch_ldw * 2) % 16 == 0 && a->prefetch_K % 8 == 0,
                     "decode_linear: the prefetched matrix must be 16-byte aligned with 16-byte row pitches");
        p.pf_w = reinterpret_cast<const uint8_t*>(a->prefetch_w); p.pf_pitch = a->prefetch_ldw * 2;
        p.pf_N = a->prefetch_N; p.pf_rowbytes = a->prefetch_K * 2;
        const int ptiles = (a->prefetch_N + DS_ROWS - 1) / DS_ROWS;
        p.pf_spt = (a->prefetch_K + DS_KW - 1) / DS_KW;
        p.pf_total = ptiles * p.pf_spt;
        p.pf_grid = h->num_sms < ptiles ? h->num_sms : ptiles;       // the successor's grid (same rule as below)
        // per-CTA budget in 16 KB stages: the caller's request, capped by the option (default 256 KB: 37 MB over 148 SMs)
        const int cap_st = (h->ds_prefetch_kb > 0 ? h->ds_prefetch_kb : 256) / (DS_W_BYTES / 1024);
        const int want = a->prefetch_stages > 0 ? a->prefetch_stages : cap_st;
        p.pf_n = h->ds_prefetch_kb > 0 ? (want < cap_st ? want : cap_st) : want;
        if (p.pf_n <= 0) p.pf_w = nullptr;
    }
    if (!(h->attr_done & (1ull << 20))) {
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(decode_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(decode_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        h->attr_done |= 1ull << 20;
    }
    if (p.resident) IVLM_CHECK_CUDA(launch_k(h, decode_stream_kernel<true>, dim3(grid), dim3(DS_THREADS), smem, stream, *tw, *ta, p));
    else IVLM_CHECK_CUDA(launch_k(h, decode_stream_kernel<false>, dim3(grid), dim3(DS_THREADS), smem, stream, *tw, *ta, p));
    h->launches++;
    return IVLM_OK;
}

extern "C" int ivlm_decode_chain(ivlm_handle h, const ivlm_decode_linear_args* phases, int32_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && phases && n >= 1 && n <= DC_MAX_PH, "decode_chain: 1..%d phases", DC_MAX_PH);
    static_assert(sizeof(ChainMaps) + sizeof(DecodeChainParams) <= 4000, "kernel parameter space");
    DecodeChainParams cp = {};
    ChainMaps maps;
    memset(&maps, 0, sizeof(maps));
    cp.n = n;
    int n_rope = 0, max_k_res = 0;
    for (int i = 0; i < n; ++i) {
        const ivlm_decode_linear_args* a = phases + i;
        IVLM_TRY(fill_decode_params(h, a, cp.ph[i]));
        cp.ph[i].resident = a->norm_gamma != nullptr ? 1 : 0;   // RMSNorm needs the whole row; everything else streams with the weights
        if (cp.ph[i].resident && a->K > max_k_res) max_k_res = a->K;
        n_rope += a->epilogue == DS_EPI_ROPE_KV;
        const CUtensorMap *tw, *ta;
        IVLM_TRY(get_tmap_bf16_kchunk3d(h, a->w, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, DS_ROWS, &tw));
        IVLM_TRY(get_tmap_bf16_kchunk3d(h, a->a, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, DS_MAX_M, &ta));
        maps.w[i] = *tw;
        maps.a[i] = *ta;
    }
    IVLM_REQUIRE(n_rope <= 1, "decode_chain: at most one ROPE_KV phase (one set of rotary tables in shared memory)");
    cp.max_k_res = max_k_res;
    const size_t cap = 227 * 1024;
    const size_t fixed = DS_FIXED_BYTES + (size_t)DS_MAX_M * (max_k_res + 8) * 2 + (size_t)max_k_res * 2 + 1024;
    IVLM_REQUIRE(fixed + 3 * (size_t)(DS_W_BYTES + DS_A_BYTES) <= cap, "decode_chain: K = %d of a phase with fused RMSNorm leaves no room for the ring", max_k_res);
    int ns = (int)((cap - fixed) / (DS_W_BYTES + DS_A_BYTES));
    if (ns > 6) ns = 6;
    if (h->ds_stages >= 3 && h->ds_stages < ns) ns = h->ds_stages;
    cp.nstages = ns;
    cp.gbar = reinterpret_cast<int*>(h->ws + DS_FLAG_OFFSET_BYTES) + 768;   // behind the per-CTA hand-over flags (<= 512 SMs)
    const size_t smem = fixed + (size_t)ns * (DS_W_BYTES + DS_A_BYTES);
    // every CTA waits for every other one: the grid must be resident at once
    int grid = h->num_sms;
    int min_tiles = cp.ph[0].tiles;
    for (int i = 1; i < n; ++i) min_tiles = cp.ph[i].tiles < min_tiles ? cp.ph[i].tiles : min_tiles;
    if (min_tiles < grid && n == 1) grid = min_tiles;
    if (!(h->attr_done & (1ull << 22))) {
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        int per_sm = 0;
        IVLM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_chain_kernel, DS_THREADS, smem));
        IVLM_REQUIRE(per_sm >= 1, "decode_chain: the kernel does not fit on an SM with %zu bytes of shared memory", smem);
        h->attr_done |= 1ull << 22;
    }
    IVLM_CHECK_CUDA(launch_k(h, decode_chain_kernel, dim3(grid), dim3(DS_THREADS), smem, stream, maps, cp));
    h->launches++;
    return IVLM_OK;
}
