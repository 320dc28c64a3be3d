// out[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual)   -- the nn.Linear contraction on the hot path
// (SAM ViT-H qkv/proj/MLP, CLIP-L, LLaMA-13B projections, mask-decoder image-side projections).
//
// Blackwell-native design (sm_100a):
//   * persistent CTAs (one per SM), static round-robin tile scheduler, optional split-K work units;
//   * warp 0 lane 0: TMA producer (cp.async.bulk.tensor, 128B swizzle, multi-stage mbarrier ring);
//   * warp 1 lane 0: tcgen05.mma issuer (UMMA 128 x BN x 16, bf16 -> fp32 accumulators in TMEM);
//   * warp 2: TMEM allocator; warps 4-11: epilogue.  bf16 row-major outputs go TMEM -> registers (bias, activation,
//     bf16 rounding; activation is a template parameter so the per-element code is branch-free and the compiler
//     interleaves independent elements) -> 128B-swizzled shared-memory slab -> coalesced 16-byte global stores
//     with the residual add and the optional row scatter;
//   * TMEM accumulators are double buffered (2 x BN columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1.
// Small token counts (LLaMA decode, mask-decoder tokens) are run "swapped": the weight matrix is the
// 128-row UMMA operand, the tokens are the narrow BN operand, and the epilogue writes transposed.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quadrant: each takes half of the columns
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int GEMM_THREADS = (EPI_WARP0 + EPI_WARPS) * 32;

struct GemmParams {
    int M, N, K;  // kernel view: A-operand rows (128-row tiles), B-operand rows (BN tiles), reduction length
    int num_m_tiles, num_n_tiles;
    int k_blocks, k_splits, k_blocks_per_split;
    void* out;
    long long out_rs, out_cs;  // element strides of out for (row r of A-operand, row c of B-operand)
    const bf16* bias;
    int bias_on_rows;  // bias indexed by r (swapped mode) instead of c
    const bf16* res;
    long long res_rs, res_cs;
    const int* row_map;  // optional remap of r -> output row (normal mode only), -1 drops the row
    int res_row_mod;     // >0: residual row = output row % res_row_mod (normal mode only)
    int act;
    int out_f32;      // out is fp32 (else bf16)
    int atomic;       // split-K: red.add.f32 into out (fp32), no bias/act/res
    int round_steps;  // mirror torch bf16 op boundaries: round after bias, after act, after residual
    int fused_split;  // split-K with in-kernel reduction: partial tiles -> workspace, the last CTA of a tile sums them
    float* ws_partial;  // [tiles * k_splits][BM][BN] fp32
    int* ws_counter;    // [tiles], zero between launches (the reducing CTA resets its counter)
    int a_static;     // operand A is a weight matrix (swapped form): safe to prefetch before griddepcontrol.wait
    int n_fastest;    // tile order (see tile_coords)
    int staged;       // bf16 row-major output through the shared-memory staged epilogue
};

// Tile order: the operand that is re-read across the other dimension should be the small one, so that it stays in
// L2.  n_fastest walks all N tiles of a few M tiles first (weights resident, activations streamed once); the default
// walks M first (activations resident, weights streamed once -- LLaMA prefill, where W >> A).
IVLM_DEVINL void tile_coords(const GemmParams& p, int t, int bn, int& m0, int& n0) {
    if (p.n_fastest) {
        n0 = (t % p.num_n_tiles) * bn;
        m0 = (t / p.num_n_tiles) * BM;
    } else {
        m0 = (t % p.num_m_tiles) * BM;
        n0 = (t / p.num_m_tiles) * bn;
    }
}

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int MAX_STAGES = 8;
    static constexpr int SMEM_BUDGET = 196608;  // 4 x 48 KB stages at BN=256; leaves room for the epilogue staging
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) < MAX_STAGES ? (SMEM_BUDGET / STAGE_BYTES) : MAX_STAGES;
    static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
    // epilogue staging: two [128 rows x 64 bf16] slabs (128B-swizzled) + one fp32 bias row
    static constexpr bool STAGED = BN >= 64;
    static constexpr int EPI_BYTES = STAGED ? (2 * BM * 128 + BN * 4) : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Phase 1 of the staged epilogue for one warp: 32 accumulator rows x 32 columns -> bf16 -> swizzled slab.
template <int ACT>
IVLM_DEVINL void epi_stage_chunk(const uint32_t (&raw)[32], const float* __restrict__ bs, bool has_bias, uint8_t* buf, int rloc,
                                 int seg0) {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        float y0 = __uint_as_float(raw[j]), y1 = __uint_as_float(raw[j + 1]);
        if (has_bias) {   // one FADD2 for the pair (bias_s is 8-byte aligned, j is even)
            const float2 b2 = *reinterpret_cast<const float2*>(bs + j);
            f32x2_unpack(f32x2_add(f32x2_pack(y0, y1), f32x2_pack(b2.x, b2.y)), y0, y1);
        }
        if (ACT != ACT_NONE) {
            bf16_round_pair(y0, y1);   // the Linear output as the reference holds it (bf16) before the activation
            if (ACT == ACT_GELU) {
                gelu_fast_pair(y0, y1);
            } else {
                y0 = apply_act_fast(y0, ACT);
                y1 = apply_act_fast(y1, ACT);
            }
        }
        x[j] = y0;
        x[j + 1] = y1;
    }
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        uint4 q;
        q.x = pack_bf16x2(x[j8 * 8 + 0], x[j8 * 8 + 1]); q.y = pack_bf16x2(x[j8 * 8 + 2], x[j8 * 8 + 3]);
        q.z = pack_bf16x2(x[j8 * 8 + 4], x[j8 * 8 + 5]); q.w = pack_bf16x2(x[j8 * 8 + 6], x[j8 * 8 + 7]);
        *reinterpret_cast<uint4*>(buf + rloc * 128 + (((seg0 + j8) ^ (rloc & 7)) << 4)) = q;
    }
}

// Phase 1 for ACT_SWIGLU: the warp's 32 accumulator columns are [gate 0-7 | up 0-7 | gate 8-15 | up 8-15] of 16 features ->
// 16 outputs bf16(bf16(silu(bf16(gate))) * bf16(up)) (the rounding points of silu_mul_kernel) -> two 16-byte segments.
IVLM_DEVINL void epi_stage_swiglu(const uint32_t (&raw)[32], uint8_t* buf, int rloc, int seg0) {
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float g0 = __uint_as_float(raw[h8 * 16 + j]), u0 = __uint_as_float(raw[h8 * 16 + 8 + j]);
            float g1 = __uint_as_float(raw[h8 * 16 + j + 1]), u1 = __uint_as_float(raw[h8 * 16 + 8 + j + 1]);
            bf16_round_pair(g0, u0);
            bf16_round_pair(g1, u1);
            float s0 = apply_act_fast(g0, ACT_SILU), s1 = apply_act_fast(g1, ACT_SILU);
            bf16_round_pair(s0, s1);
            o[j] = s0 * u0;
            o[j + 1] = s1 * u1;
        }
        uint4 q;
        q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]); q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(buf + rloc * 128 + (((seg0 + h8) ^ (rloc & 7)) << 4)) = q;
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzle atoms need 1024-byte aligned tiles.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;                       // 2 x [128][128 B]
    float* bias_s = reinterpret_cast<float*>(epi_stage + 2 * BM * 128);         // [BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    volatile int* split_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_units = p.num_m_tiles * p.num_n_tiles * p.k_splits;
    pdl_launch();  // the successor's launch + prologue (and weight prefetch, if it is a swapped GEMM) overlap this kernel

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // Programmatic dependent launch: operand A of the swapped form is a static weight matrix, so the first
            // pipeline stages of this CTA's first unit can start streaming before the predecessor kernel has finished.
            int prefetched = 0;
            if (p.a_static && blockIdx.x < total_units) {
                const int u = blockIdx.x;
                const int ks = u % p.k_splits;
                int m0, n0;
                tile_coords(p, u / p.k_splits, BN, m0, n0);
                const int kb0 = ks * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks);
                for (int kb = kb0; kb < kb1 && prefetched < STAGES; ++kb, ++prefetched) {
                    mbar_arrive_expect_tx(&full_bar[prefetched], Cfg::STAGE_BYTES);
                    tma_load_2d(smem_a + prefetched * Cfg::A_BYTES, &tmA, &full_bar[prefetched], kb * BK, m0);
                }
            }
            pdl_wait();
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const int ks = u % p.k_splits;
                const int t = u / p.k_splits;
                int m0, n0;
                tile_coords(p, t, BN, m0, n0);
                const int kb0 = ks * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (prefetched > 0) {  // weights of this stage are already in flight: add the activation tile
                        --prefetched;
                        tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
                    } else {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, m0);
                        tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++it) {
                const int ks = u % p.k_splits;
                const int kb0 = ks * p.k_blocks_per_split;
                const int kb1 = min(kb0 + p.k_blocks_per_split, p.k_blocks);
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ------------------------------------------------------------ epilogue
        pdl_wait();  // bias / residual / row_map / split-K workspace may come from the predecessor
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
        const int chalf = (warp - EPI_WARP0) >> 2;  // which half of the columns this warp of the quadrant takes
        int it = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++it) {
            const int t = u / p.k_splits;
            int m0, n0;
            tile_coords(p, t, BN, m0, n0);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();

            if constexpr (Cfg::STAGED) {
                if (p.staged) {
                    // ---- staged epilogue.  Phase 1: warp (quad, chalf) converts its 32 rows x 32 columns of the
                    // 64-column slab into the swizzled staging buffer; phase 2: every thread moves (row, 16-byte
                    // segment) items to global memory, so stores and residual loads are whole 128-byte row pieces.
                    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
                    if (p.bias != nullptr)
                        for (int c = et; c < BN; c += EPI_THREADS)
                            bias_s[c] = (n0 + c < p.N) ? __bfloat162float(p.bias[n0 + c]) : 0.f;
                    named_bar_sync(1, EPI_THREADS);
                    constexpr int NSLAB = BN / 64;
                    const int rloc = quad * 32 + lane;
                    const bool has_bias = p.bias != nullptr;
                    // the TMEM read of slab s+1 is issued as soon as slab s has been converted, so its latency runs under
                    // the barrier and the global stores of slab s
                    const uint32_t taddr0 = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(as * BN + chalf * 32);
                    uint32_t raw[32];
                    tmem_ld_32x32(taddr0, raw);
#pragma unroll 1
                    for (int slab = 0; slab < NSLAB; ++slab) {
                        uint8_t* buf = epi_stage + (slab & 1) * (BM * 128);
                        const float* bs = bias_s + slab * 64 + chalf * 32;
                        tmem_ld_wait();
                        switch (p.act) {
                            case ACT_GELU: epi_stage_chunk<ACT_GELU>(raw, bs, has_bias, buf, rloc, chalf * 4); break;
                            case ACT_QUICK_GELU: epi_stage_chunk<ACT_QUICK_GELU>(raw, bs, has_bias, buf, rloc, chalf * 4); break;
                            case ACT_RELU: epi_stage_chunk<ACT_RELU>(raw, bs, has_bias, buf, rloc, chalf * 4); break;
                            case ACT_SILU: epi_stage_chunk<ACT_SILU>(raw, bs, has_bias, buf, rloc, chalf * 4); break;
                            case ACT_SWIGLU: epi_stage_swiglu(raw, buf, rloc, chalf * 2); break;
                            default: epi_stage_chunk<ACT_NONE>(raw, bs, has_bias, buf, rloc, chalf * 4); break;
                        }
                        if (slab + 1 < NSLAB) {
                            tmem_ld_32x32(taddr0 + uint32_t((slab + 1) * 64), raw);
                        } else {  // accumulator fully read: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[as]);
                        }
                        named_bar_sync(1, EPI_THREADS);
                        const int seg = et & 7;
                        // SWIGLU: a 64-column slab of accumulators is 32 output columns (segments 0-3 of the staging rows)
                        const bool sw = p.act == ACT_SWIGLU;
                        const int col = sw ? (n0 >> 1) + slab * 32 + seg * 8 : n0 + slab * 64 + seg * 8;
                        const int ncol = (sw && seg >= 4) ? 0 : (sw ? (p.N >> 1) : p.N);
                        // all row-map lookups first, then all residual loads, then add + store: the loads of the four rows a
                        // thread moves are in flight together instead of one dependent chain per row
                        constexpr int NIT = BM * 8 / EPI_THREADS;
                        int orow_[NIT];
                        uint4 r4_[NIT];
#pragma unroll
                        for (int i = 0; i < NIT; ++i) {
                            const int grow = m0 + i * (EPI_THREADS / 8) + (et >> 3);
                            int orow = (grow < p.M && col < ncol) ? grow : -1;
                            if (orow >= 0 && p.row_map != nullptr) orow = p.row_map[grow];
                            orow_[i] = orow;
                        }
                        if (p.res != nullptr) {
#pragma unroll
                            for (int i = 0; i < NIT; ++i) {
                                if (orow_[i] < 0) continue;
                                const int rrow = p.res_row_mod > 0 ? (orow_[i] % p.res_row_mod) : orow_[i];
                                r4_[i] = *reinterpret_cast<const uint4*>(p.res + (long long)rrow * p.res_rs + col);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < NIT; ++i) {
                            if (orow_[i] < 0) continue;
                            const int rl = i * (EPI_THREADS / 8) + (et >> 3);
                            uint4 q = *reinterpret_cast<const uint4*>(buf + rl * 128 + ((seg ^ (rl & 7)) << 4));
                            if (p.res != nullptr) {
                                q.x = add_bf16x2(q.x, r4_[i].x); q.y = add_bf16x2(q.y, r4_[i].y);
                                q.z = add_bf16x2(q.z, r4_[i].z); q.w = add_bf16x2(q.w, r4_[i].w);
                            }
                            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (long long)orow_[i] * p.out_rs + col) = q;
                        }
                    }
                    continue;
                }
            }

            const int r = m0 + quad * 32 + lane;
            const bool r_ok = r < p.M;
            int orow = r;
            if (p.row_map != nullptr && r_ok) orow = p.row_map[r];
            const bool row_live = r_ok && orow >= 0;
            float bias_r = 0.f;
            if (p.bias != nullptr && p.bias_on_rows && r_ok) bias_r = __bfloat162float(p.bias[r]);

            constexpr int CH = BN < 32 ? BN : 32;  // columns per TMEM load
            // bias -> act -> residual -> store for one chunk of CH accumulator columns held in v
            auto finish_chunk = [&](float (&v)[CH], int c0) {
                const int cbase = n0 + c0;
                if (!row_live || cbase >= p.N) return;
                if (p.atomic) {
                    float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const int c = cbase + j;
                        if (c < p.N) atomicAdd(o + (long long)orow * p.out_rs + (long long)c * p.out_cs, v[j]);
                    }
                    return;
                }
                // bias -> act -> residual, rounding like the eager bf16 reference at each op boundary
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int c = cbase + j;
                    float x = v[j];
                    if (p.bias != nullptr) {
                        x += p.bias_on_rows ? bias_r : ((c < p.N) ? __bfloat162float(p.bias[c]) : 0.f);
                    }
                    if (p.round_steps) x = bf16_round(x);
                    if (p.act != ACT_NONE) {
                        x = apply_act(x, p.act);
                        if (p.round_steps) x = bf16_round(x);
                    }
                    v[j] = x;
                }
                if (p.out_cs == 1) {
                    // row-major output: this thread owns CH consecutive columns of one row
                    const long long obase = (long long)orow * p.out_rs + cbase;
                    if (p.res != nullptr) {
                        const int rrow = p.res_row_mod > 0 ? (orow % p.res_row_mod) : orow;
                        const bf16* rp = p.res + (long long)rrow * p.res_rs + cbase;
                        if (cbase + CH <= p.N) {
#pragma unroll
                            for (int j = 0; j < CH; j += 8) {
                                uint4 q = *reinterpret_cast<const uint4*>(rp + j);
                                float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z),
                                       f3 = unpack_bf16x2(q.w);
                                v[j + 0] += f0.x; v[j + 1] += f0.y; v[j + 2] += f1.x; v[j + 3] += f1.y;
                                v[j + 4] += f2.x; v[j + 5] += f2.y; v[j + 6] += f3.x; v[j + 7] += f3.y;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (cbase + j < p.N) v[j] += __bfloat162float(rp[j]);
                        }
                    }
                    if (p.out_f32) {
                        float* o = reinterpret_cast<float*>(p.out) + obase;
                        if (cbase + CH <= p.N) {
#pragma unroll
                            for (int j = 0; j < CH; j += 4)
                                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (cbase + j < p.N) o[j] = v[j];
                        }
                    } else {
                        bf16* o = reinterpret_cast<bf16*>(p.out) + obase;
                        if (cbase + CH <= p.N) {
#pragma unroll
                            for (int j = 0; j < CH; j += 8) {
                                uint4 q;
                                q.x = pack_bf16x2(v[j + 0], v[j + 1]);
                                q.y = pack_bf16x2(v[j + 2], v[j + 3]);
                                q.z = pack_bf16x2(v[j + 4], v[j + 5]);
                                q.w = pack_bf16x2(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(o + j) = q;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (cbase + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
                        }
                    }
                } else {
                    // transposed output (swapped operands): lanes hold consecutive r -> coalesced per column
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const int c = cbase + j;
                        if (c >= p.N) continue;
                        float x = v[j];
                        if (p.res != nullptr)
                            x += __bfloat162float(p.res[(long long)orow * p.res_rs + (long long)c * p.res_cs]);
                        const long long oi = (long long)orow * p.out_rs + (long long)c * p.out_cs;
                        if (p.out_f32) reinterpret_cast<float*>(p.out)[oi] = x;
                        else reinterpret_cast<bf16*>(p.out)[oi] = __float2bfloat16_rn(x);
                    }
                }
            };
            const int ks_unit = u % p.k_splits;
            float* my_part = p.ws_partial + ((size_t)(t * p.k_splits + ks_unit) * BM + quad * 32 + lane) * BN;
#pragma unroll 1
            for (int c0 = chalf * CH; c0 < BN; c0 += 2 * CH) {
                float v[CH];
                {
                    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(as * BN + c0);
                    if constexpr (CH == 32) {
                        uint32_t raw[32];
                        tmem_ld_32x32(taddr, raw);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    } else {
                        uint32_t raw[16];
                        tmem_ld_32x16(taddr, raw);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                    }
                }
                if (p.fused_split) {
                    // raw fp32 partial of this K split (64 contiguous bytes or more per thread)
#pragma unroll
                    for (int j = 0; j < CH; j += 4)
                        *reinterpret_cast<float4*>(my_part + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    continue;
                }
                finish_chunk(v, c0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (p.fused_split) {
                // publish the partial, count arrivals; the CTA that completes the tile reduces all splits in split
                // order (deterministic) and runs the normal epilogue on the sum
                __threadfence();
                named_bar_sync(1, EPI_THREADS);
                if (threadIdx.x == EPI_WARP0 * 32) {
                    const int old = atomicAdd(p.ws_counter + t, 1);
                    const int last = old == p.k_splits - 1;
                    if (last) p.ws_counter[t] = 0;
                    *split_flag = last;
                }
                named_bar_sync(1, EPI_THREADS);
                if (*split_flag) {
                    __threadfence();
                    const float* base = p.ws_partial + ((size_t)t * p.k_splits * BM + quad * 32 + lane) * BN;
#pragma unroll 1
                    for (int c0 = chalf * CH; c0 < BN; c0 += 2 * CH) {
                        float v[CH];
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = 0.f;
                        for (int sp = 0; sp < p.k_splits; ++sp) {
                            const float* src = base + (size_t)sp * BM * BN + c0;
#pragma unroll
                            for (int j = 0; j < CH; j += 4) {
                                const float4 f = __ldcg(reinterpret_cast<const float4*>(src + j));
                                v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
                            }
                        }
                        finish_chunk(v, c0);
                    }
                }
                named_bar_sync(1, EPI_THREADS);  // split_flag is reused by the next unit
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- host side

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

int get_tmap_bf16_ex(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols, uint32_t swizzle_bytes, const CUtensorMap** out) {
    TmapKey key{ptr, rows, cols, ld, box_rows, box_cols, swizzle_bytes};
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
        *out = &it->second;
        return IVLM_OK;
    }
    it = h->tmaps_old.find(key);
    if (it != h->tmaps_old.end()) {
        *out = &it->second;
        return IVLM_OK;
    }
    auto enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return IVLM_ERR_CUDA;
    }
    IVLM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand base %p is not 16-byte aligned", ptr);
    IVLM_REQUIRE((ld * 2) % 16 == 0, "TMA operand row pitch %llu elements is not a multiple of 8",
                 (unsigned long long)ld);
    IVLM_REQUIRE(box_cols * 2 <= swizzle_bytes && box_rows <= 256, "TMA box %ux%u does not fit swizzle %u", box_rows,
                 box_cols, swizzle_bytes);
    const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                         : CU_TENSOR_MAP_SWIZZLE_32B;
    CUtensorMap m;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u swizzle=%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols,
                  swizzle_bytes);
        return IVLM_ERR_CUDA;
    }
    if (h->tmaps.size() >= IVLM_TMAP_GEN) {  // retire a generation (see runtime.h): earlier pointers stay valid
        h->tmaps_old.swap(h->tmaps);
        h->tmaps.clear();
    }
    auto ins = h->tmaps.emplace(key, m);
    *out = &ins.first->second;
    return IVLM_OK;
}

// 3-D view of a row-major [rows, K] bf16 matrix for the weight-streaming decode kernel: (64-element k-chunk, row, chunk index),
// box (64, box_rows, 8), 128-byte swizzle.  One box = 8 K-major slabs of [box_rows x 128 B] (512 columns of box_rows rows).
int get_tmap_bf16_kchunk3d(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows,
                           const CUtensorMap** out) {
    TmapKey key{ptr, rows, K, ld, box_rows, 64, 128 + 3};   // swizzle tag 131: rank-3 k-chunk map
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) { *out = &it->second; return IVLM_OK; }
    it = h->tmaps_old.find(key);
    if (it != h->tmaps_old.end()) { *out = &it->second; return IVLM_OK; }
    auto enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return IVLM_ERR_CUDA;
    }
    IVLM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0 && K % 64 == 0 && box_rows <= 256,
                 "k-chunk tensor map: base/pitch must be 16-byte aligned and K a multiple of 64 (K=%llu)", (unsigned long long)K);
    CUtensorMap m;
    cuuint64_t gdim[3] = {64, rows, K / 64};
    cuuint64_t gstr[2] = {ld * 2, 128};
    cuuint32_t box[3] = {64, box_rows, 8};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (k-chunk 3-D) failed (%d) rows=%llu K=%llu ld=%llu", (int)r, (unsigned long long)rows,
                  (unsigned long long)K, (unsigned long long)ld);
        return IVLM_ERR_CUDA;
    }
    if (h->tmaps.size() >= IVLM_TMAP_GEN) {
        h->tmaps_old.swap(h->tmaps);
        h->tmaps.clear();
    }
    auto ins = h->tmaps.emplace(key, m);
    *out = &ins.first->second;
    return IVLM_OK;
}

int get_tmap_bf16(ivlm_ctx* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  const CUtensorMap** out) {
    return get_tmap_bf16_ex(h, ptr, rows, cols, ld, box_rows, BK, 128, out);
}

template <int BN>
static int launch_gemm(ivlm_ctx* h, const CUtensorMap* ta, const CUtensorMap* tb, const GemmParams& p,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    // the attribute is per device: remembered per handle (a handle is bound to one device), bit = log2(BN)
    const uint64_t attr_bit = 1ull << (BN == 256 ? 8 : BN == 128 ? 7 : BN == 64 ? 6 : BN == 32 ? 5 : 4);
    if (!(h->attr_done & attr_bit)) {
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES));
        h->attr_done |= attr_bit;
    }
    const int units = p.num_m_tiles * p.num_n_tiles * p.k_splits;
    // SM partitioning (option "sm_limit"): a token-major GEMM of a low-priority stream keeps to a subset of the SMs so
    // that the weight-streaming decode chain of another stream always finds free ones
    const int cap = (h->sm_limit > 0 && h->sm_limit < h->num_sms && !p.a_static) ? h->sm_limit : h->num_sms;
    const int grid = units < cap ? units : cap;
    IVLM_CHECK_CUDA(launch_k(h, gemm_bf16_tcgen05_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, *ta, *tb, p));
    h->launches++;
    return IVLM_OK;
}

static int pick_bn(int n) {
    if (n > 128) return 256;
    if (n > 64) return 128;
    if (n > 32) return 64;
    if (n > 16) return 32;
    return 16;
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_gemm_bf16(ivlm_handle h, const ivlm_gemm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && a, "null handle/args");
    IVLM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
    IVLM_REQUIRE(a->K % 8 == 0, "gemm: K=%d must be a multiple of 8", a->K);
    // k_splits > 1: caller-managed split-K (atomic fp32 accumulation into a pre-zeroed `out`, finalize separately).
    // k_splits == 0/1 with a workspace bound: the library may split K itself (fused, deterministic) for
    // weight-streaming shapes that would otherwise occupy only a fraction of the SMs.
    const bool split = a->k_splits > 1;
    IVLM_REQUIRE(!split || (a->out_dtype == IVLM_F32 && !a->bias && !a->residual && a->act == 0),
                 "gemm: split-K accumulates raw fp32 (no bias/act/residual)");

    IVLM_REQUIRE(a->act != ACT_SWIGLU || (a->M > 64 && a->force_swap <= 0 && !split),
                 "gemm: SWIGLU needs more than 64 tokens (the decode path fuses it in ivlm_decode_linear)");
    // Small token counts are weight streaming (HBM-bound): dedicated kernel built around the weight stream.
    if (h->small_m_variant == 0 && a->M <= h->gv_max_m && a->N <= h->gv_max_n && a->force_swap >= 0 && a->row_map == nullptr && a->k_splits <= 1 &&
        a->K % 32 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0 && a->res_row_mod == 0 &&
        (reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0)
        return launch_gemv_small_m(h, a, stream);
    // Swap operands when the token count is small: the weight rows fill the 128-row UMMA operand.
    const bool swap = a->force_swap == 1 || (a->force_swap == 0 && a->M <= 64 && a->row_map == nullptr);
    GemmParams p{};
    const void *pa, *pb;
    int64_t lda, ldb;
    if (!swap) {
        p.M = a->M; p.N = a->N;
        pa = a->a; lda = a->lda; pb = a->w; ldb = a->ldw;
        p.out_rs = a->ldo; p.out_cs = 1;
        p.res_rs = a->ldr; p.res_cs = 1;
        p.bias_on_rows = 0;
        IVLM_REQUIRE(a->N % 8 == 0 && a->ldo % 8 == 0, "gemm: row-major epilogue needs N, ldo multiples of 8");
        IVLM_REQUIRE(!a->residual || a->ldr % 8 == 0, "gemm: residual pitch must be a multiple of 8");
    } else {
        p.M = a->N; p.N = a->M;
        pa = a->w; lda = a->ldw; pb = a->a; ldb = a->lda;
        p.out_rs = 1; p.out_cs = a->ldo;
        p.res_rs = 1; p.res_cs = a->ldr;
        p.bias_on_rows = 1;
        IVLM_REQUIRE(a->row_map == nullptr, "gemm: row_map unsupported with swapped operands");
    }
    p.K = a->K;
    const int bn = pick_bn(p.N);
    p.num_m_tiles = (p.M + BM - 1) / BM;
    p.num_n_tiles = (p.N + bn - 1) / bn;
    p.k_blocks = (p.K + BK - 1) / BK;
    p.k_splits = split ? a->k_splits : 1;
    p.fused_split = 0;
    if (!split && swap && h->ws != nullptr && a->k_splits == 0) {
        const int tiles = p.num_m_tiles * p.num_n_tiles;
        // cost in k-block units: waves x (k-blocks per unit + pipeline fill / drain / partial-tile exchange per unit).  The
        // exchange grows with the token tile: ~6 k-blocks at 16 tokens, 4x / 16x that at 32 / 64 (measured with
        // tools/prof_decode.py splits: at 32 / 64 tokens splitting gate-up in two costs 73 / 87 us against 60 / 61 unsplit,
        // while the 40-tile o / down projections still gain from a 3-way split)
        int best = 1;
        const double per_unit = bn <= 16 ? 6.0 : 6.0 * (bn / 16.0) * (bn / 16.0);
        auto cost_of = [&](int sp) {
            return (double)((tiles * sp + h->num_sms - 1) / h->num_sms) * ((double)p.k_blocks / sp + per_unit);
        };
        double best_cost = cost_of(1);
        for (int sp = 2; sp <= 8; ++sp) {
            if (p.k_blocks / sp < 8) break;
            const double cost = cost_of(sp);
            if (cost < best_cost * 0.9) { best_cost = cost; best = sp; }
        }
        if (h->fused_split_force > 0) best = h->fused_split_force;  // A/B knob (tools/prof_decode.py splits)
        const size_t need = (size_t)tiles * best * BM * bn * sizeof(float) + (size_t)tiles * sizeof(int) + 256;
        if (best > 1 && need <= h->ws_bytes - IVLM_WS_COUNTER_BYTES && (size_t)tiles * sizeof(int) <= IVLM_WS_COUNTER_BYTES - 4096 /* the last 4 KB are decode_stream's flags */) {
            p.k_splits = best;
            p.fused_split = 1;
            p.ws_counter = reinterpret_cast<int*>(h->ws);
            p.ws_partial = reinterpret_cast<float*>(h->ws + IVLM_WS_COUNTER_BYTES);
        }
    }
    if (p.k_splits > p.k_blocks) p.k_splits = p.k_blocks;
    p.k_blocks_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
    p.k_splits = (p.k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;  // no empty splits
    p.out = a->out;
    p.bias = reinterpret_cast<const bf16*>(a->bias);
    p.res = reinterpret_cast<const bf16*>(a->residual);
    p.row_map = a->row_map;
    p.res_row_mod = a->res_row_mod;
    IVLM_REQUIRE(!(swap && a->res_row_mod > 0), "gemm: res_row_mod unsupported with swapped operands");
    p.act = a->act;
    p.out_f32 = a->out_dtype == IVLM_F32;
    p.atomic = split ? 1 : 0;
    p.round_steps = (a->out_dtype == IVLM_BF16 && !a->no_round) ? 1 : 0;
    p.staged = (!swap && !split && a->out_dtype == IVLM_BF16 && !a->no_round && bn >= 64) ? 1 : 0;
    if (a->act == ACT_SWIGLU)
        IVLM_REQUIRE(p.staged && a->N % 16 == 0 && !a->bias && !a->residual && a->row_map == nullptr,
                     "gemm: SWIGLU runs in the staged epilogue only (token count > 64, N >= 64 and a multiple of 16, bf16 output, no bias / "
                     "residual / row_map); got M=%d N=%d", a->M, a->N);
    // keep the smaller operand L2-resident across the sweep of the other dimension
    p.n_fastest = ((long long)p.N * p.K < (long long)p.M * p.K) ? 1 : 0;
    p.a_static = swap ? 1 : 0;

    const CUtensorMap *ta, *tb;
    IVLM_TRY(get_tmap_bf16(h, pa, p.M, p.K, lda, BM, &ta));
    IVLM_TRY(get_tmap_bf16(h, pb, p.N, p.K, ldb, bn, &tb));
    switch (bn) {
        case 256: return launch_gemm<256>(h, ta, tb, p, stream);
        case 128: return launch_gemm<128>(h, ta, tb, p, stream);
        case 64: return launch_gemm<64>(h, ta, tb, p, stream);
        case 32: return launch_gemm<32>(h, ta, tb, p, stream);
        default: return launch_gemm<16>(h, ta, tb, p, stream);
    }
}
