// SAM ViT attention (head_dim 80, decomposed relative-position bias) on the 5th-gen tensor cores.
//
//   out = softmax(q k^T * scale + rel_h[q, ky] + rel_w[q, kx]) v          image_encoder.py:235-260, :354-392
//
// One CTA per (128-query tile, head, image-or-window).  Everything GEMM-shaped runs as tcgen05.mma with fp32
// accumulators in TMEM, operands staged by TMA:
//   prologue  T_h = Q Rh_all^T, T_w = Q Rw_all^T   (the two rel-pos einsums for the whole tile: 2 x 5 UMMAs instead of a
//             separate kernel and a 1 GB/block fp32 round trip through HBM); rounded to bf16 like the reference's einsum
//             outputs and kept in shared memory, transposed ([table index][query row]) so that lookups are conflict-free;
//   per 128-key tile   S = Q K^T (5 UMMAs, K = 64 + 16) -> softmax warps read S with tcgen05.ld (one query row per
//             thread, no shuffles), add the bias, online softmax with lazy rescaling (O in TMEM is only touched when the
//             running max moves by more than 2^8), P (bf16) -> 128B-swizzled shared memory -> O += P V (V is the
//             MN-major operand straight from its row-major [key, 80] layout; N = 64 + 16).
// head_dim 80 is split 64 + 16: the 64-column part uses 128B-swizzled tiles, the 16-column part 32B-swizzled tiles.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = softmax / correction / epilogue.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int AT_BQ = 128, AT_BK = 128, AT_HD = 80, AT_NS = 2, AT_THREADS = 256;
constexpr int AT_QA = AT_BQ * 128, AT_QB = AT_BQ * 32;                // Q tile: 64-col part, 16-col part (bytes)
constexpr int AT_STAGE = 2 * (AT_BK * 128 + AT_BK * 32);              // K (a,b) + V (a,b)
constexpr int AT_P = AT_BQ * AT_BK * 2;                               // P tile, two 64-key chunks
constexpr int AT_TAB = 128 * AT_BQ * 2;                               // transposed bias table [<=128 idx][128 rows] bf16
constexpr int AT_SMEM = AT_QA + AT_QB + AT_NS * AT_STAGE + AT_P + 2 * AT_TAB + 1024 + 256;
constexpr int AT_TMEM_COLS = 512;
constexpr int AT_O_COL = 256;
constexpr float AT_LOG2E = 1.4426950408889634f;

struct SamAttnParams {
    bf16* out;
    long long out_ld;  // elements between consecutive tokens of `out`
    int S;             // tokens per image / window
    int heads;
    int KH;            // key grid height (== query grid height)
    float scale_log2;  // head_dim^-0.5 * log2(e)
    const int* out_map;  // window kernel only, optional: output row of input row r (window_unpartition fused), -1 drops the row
    int prefetch_ahead;  // window kernel: distance (in CTAs of the launch order) of the L2 prefetch, 0 = off
};

// K-major 32B-swizzled operand ([rows][16 bf16], rows 32 B apart, 8-row groups 256 B apart).
IVLM_DEVINL uint64_t umma_desc_sw32_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
// MN-major operand whose rows are K (keys): [k rows][64 bf16] 128B-swizzled; 8-row groups 1024 B apart (SBO).
IVLM_DEVINL uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024 >> 4) << 16;  // LBO: next 64-element MN block (unused: one block)
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major, [k rows][16 bf16] 32B-swizzled; 8-row groups 256 B apart.
IVLM_DEVINL uint64_t umma_desc_sw32_mnmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(256 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) {  // B operand MN-major
    return umma_idesc_bf16(M, N) | (1u << 16);
}

IVLM_DEVINL void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
IVLM_DEVINL void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
IVLM_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
IVLM_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// S / T tile (128 x 128 fp32 in TMEM) = A[128 x 80] . B[128 x 80]^T, both K-major, split 64 + 16.
IVLM_DEVINL void issue_qk(uint32_t d_tmem, const uint8_t* a64, const uint8_t* a16, const uint8_t* b64, const uint8_t* b16) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(a64)), db = umma_desc_sw128_kmajor(smem_u32(b64));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_bf16(d_tmem, umma_desc_sw32_kmajor(smem_u32(a16)), umma_desc_sw32_kmajor(smem_u32(b16)), idesc, 1u);
}

// KW = key grid width: 64 (global attention over the 64x64 token grid) or 14 (14x14 windows).
template <int KW>
__global__ void __launch_bounds__(AT_THREADS, 1)
sam_attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKVa, const __grid_constant__ CUtensorMap tmQKVb,
                        const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                        const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                        const SamAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_a = smem;
    uint8_t* q_b = q_a + AT_QA;
    uint8_t* stage0 = q_b + AT_QB;
    auto k_a = [&](int s) { return stage0 + s * AT_STAGE; };
    auto k_b = [&](int s) { return stage0 + s * AT_STAGE + AT_BK * 128; };
    auto v_a = [&](int s) { return stage0 + s * AT_STAGE + AT_BK * 160; };
    auto v_b = [&](int s) { return stage0 + s * AT_STAGE + AT_BK * 288; };
    uint8_t* p_s = stage0 + AT_NS * AT_STAGE;
    bf16* th_s = reinterpret_cast<bf16*>(p_s + AT_P);  // [idx][128 rows]
    bf16* tw_s = th_s + 128 * AT_BQ;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tw_s) + AT_TAB);
    uint64_t* q_full = bars;            // 1
    uint64_t* tab_full = bars + 1;      // 1
    uint64_t* t_full = bars + 2;        // 1
    uint64_t* kv_full = bars + 3;       // AT_NS
    uint64_t* kv_empty = bars + 5;      // AT_NS
    uint64_t* s_full = bars + 7;        // 2
    uint64_t* s_empty = bars + 9;       // 2
    uint64_t* p_full = bars + 11;       // 1
    uint64_t* pv_done = bars + 12;      // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int E = p.heads * AT_HD;
    const int n_tiles = (p.S + AT_BK - 1) / AT_BK;
    const int row_base = b * p.S;  // first token row of this image / window in the packed qkv matrix

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKVa); tma_prefetch_desc(&tmQKVb);
        tma_prefetch_desc(&tmRHa); tma_prefetch_desc(&tmRHb);
        tma_prefetch_desc(&tmRWa); tma_prefetch_desc(&tmRWb);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(tab_full, 1); mbar_init(t_full, 1);
        for (int s = 0; s < AT_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, AT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, AT_QA + AT_QB);
            tma_load_2d(q_a, &tmQKVa, q_full, h * AT_HD, row_base + q0);
            tma_load_2d(q_b, &tmQKVb, q_full, h * AT_HD + 64, row_base + q0);
            // rel-pos tables ride in the K slots of stages 0 and 1 until the prologue MMAs have consumed them
            mbar_arrive_expect_tx(tab_full, 2 * (AT_BK * 128 + AT_BK * 32));
            tma_load_2d(k_a(0), &tmRHa, tab_full, 0, 0);
            tma_load_2d(k_b(0), &tmRHb, tab_full, 64, 0);
            tma_load_2d(k_a(1), &tmRWa, tab_full, 0, 0);
            tma_load_2d(k_b(1), &tmRWb, tab_full, 64, 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % AT_NS;
                mbar_wait(&kv_empty[s], (j / AT_NS) & 1);  // completion 0 = prologue MMAs done with the tables
                mbar_arrive_expect_tx(&kv_full[s], AT_STAGE);
                const int r = row_base + j * AT_BK;
                tma_load_2d(k_a(s), &tmQKVa, &kv_full[s], E + h * AT_HD, r);
                tma_load_2d(k_b(s), &tmQKVb, &kv_full[s], E + h * AT_HD + 64, r);
                tma_load_2d(v_a(s), &tmQKVa, &kv_full[s], 2 * E + h * AT_HD, r);
                tma_load_2d(v_b(s), &tmQKVb, &kv_full[s], 2 * E + h * AT_HD + 64, r);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            mbar_wait(q_full, 0);
            mbar_wait(tab_full, 0);
            tc_fence_after();
            issue_qk(tmem_base + 0, q_a, q_b, k_a(0), k_b(0));        // T_h -> S buffer 0
            issue_qk(tmem_base + AT_BK, q_a, q_b, k_a(1), k_b(1));    // T_w -> S buffer 1
            umma_commit(t_full);
            for (int s = 0; s < AT_NS; ++s) umma_commit(&kv_empty[s]);
            auto issue_s = [&](int j) {
                const int s = j % AT_NS, sb = j & 1;
                mbar_wait(&kv_full[s], (j / AT_NS) & 1);
                mbar_wait(&s_empty[sb], (j >> 1) & 1);  // completion 0 = the prologue's read of T_h / T_w
                tc_fence_after();
                issue_qk(tmem_base + sb * AT_BK, q_a, q_b, k_a(s), k_b(s));
                umma_commit(&s_full[sb]);
            };
            issue_s(0);
            constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                const int s = j % AT_NS;
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < AT_BK / 16; ++ks) {
                    const uint64_t dp = umma_desc_sw128_kmajor(smem_u32(p_s + (ks >> 2) * (AT_BQ * 128))) + 2 * (ks & 3);
                    const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                    umma_bf16(tmem_base + AT_O_COL, dp, umma_desc_sw128_mnmajor(smem_u32(v_a(s) + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + AT_O_COL + 64, dp, umma_desc_sw32_mnmajor(smem_u32(v_b(s) + ks * 512)), idesc16, acc);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax / correction / epilogue
        const int quad = warp & 3;
        const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const int qtok = q0 + r;
        const int qy = qtok / KW, qx = qtok - qy * KW;
        const int KH = p.KH;
        // ---- prologue: bias tables for this tile
        mbar_wait(t_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            bf16* dst = t == 0 ? th_s : tw_s;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(lane_addr + t * AT_BK + c0, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) dst[(c0 + i) * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(raw[i]));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&s_empty[0]); mbar_arrive(&s_empty[1]); }
        // each thread only ever reads column r of the tables, which it wrote itself: no CTA-level sync needed
        const bf16* th_r = th_s + (qy + KH - 1) * AT_BQ + r;  // rel_h[ky] = th_r[-ky * 128]
        const bf16* tw_r = tw_s + (qx + KW - 1) * AT_BQ + r;  // rel_w[kx] = tw_r[-kx * 128]

        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t xr[AT_BK];  // tcgen05.ld results: not to be touched before tcgen05.wait::ld
#pragma unroll
            for (int c0 = 0; c0 < AT_BK; c0 += 32)
                tmem_ld_32x32(lane_addr + sb * AT_BK + c0, reinterpret_cast<uint32_t(&)[32]>(xr[c0]));
            tmem_ld_wait();
            float x[AT_BK];
#pragma unroll
            for (int c = 0; c < AT_BK; ++c) x[c] = __uint_as_float(xr[c]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            // ---- scale + decomposed rel-pos bias, in the log2 domain
            float mx = -INFINITY;
            if constexpr (KW == 64) {
                // tile j covers key rows 2j and 2j+1 completely; rel_w is shared by both
                const float rh0 = __bfloat162float(th_r[-(2 * j) * AT_BQ]) * AT_LOG2E;
                const float rh1 = __bfloat162float(th_r[-(2 * j + 1) * AT_BQ]) * AT_LOG2E;
#pragma unroll
                for (int kx = 0; kx < 64; ++kx) {
                    const float rw = __bfloat162float(tw_r[-kx * AT_BQ]) * AT_LOG2E;
                    x[kx] = fmaf(x[kx], p.scale_log2, rw + rh0);
                    x[64 + kx] = fmaf(x[64 + kx], p.scale_log2, rw + rh1);
                    mx = fmaxf(mx, fmaxf(x[kx], x[64 + kx]));
                }
            } else {
                const int kbase = j * AT_BK;
#pragma unroll
                for (int c = 0; c < AT_BK; ++c) {
                    const int k = kbase + c;
                    const int ky = k / KW, kx = k - ky * KW;
                    if (k < p.S) {
                        const float bias = __bfloat162float(th_r[-ky * AT_BQ]) + __bfloat162float(tw_r[-kx * AT_BQ]);
                        x[c] = fmaf(x[c], p.scale_log2, bias * AT_LOG2E);
                        mx = fmaxf(mx, x[c]);
                    } else {
                        x[c] = -INFINITY;
                    }
                }
            }
            // ---- online softmax with lazy rescaling: the reference point m_ref only moves when the row max grew by > 2^8
            float corr = 1.f;
            bool moved = false;
            if (j == 0) {
                m_ref = mx;
            } else if (mx > m_ref + 8.f) {
                corr = ex2_approx(m_ref - mx);
                m_ref = mx;
                moved = true;
            }
            float sum = 0.f;
            uint32_t pk[AT_BK / 2];
#pragma unroll
            for (int c = 0; c < AT_BK; c += 2) {
                const float p0 = ex2_approx(x[c] - m_ref), p1 = ex2_approx(x[c + 1] - m_ref);
                sum += p0 + p1;
                pk[c >> 1] = pack_bf16x2(p0, p1);
            }
            l_run = l_run * corr + sum;
            // ---- P and (rarely) the O correction may only touch shared / tensor memory once PV_{j-1} has retired
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, moved)) {
#pragma unroll
                    for (int c0 = 0; c0 < AT_HD; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld_32x16(lane_addr + AT_O_COL + c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st_32x16(lane_addr + AT_O_COL + c0, o);
                    }
                    tmem_st_wait();
                }
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint8_t* row = p_s + ch * (AT_BQ * 128) + r * 128;
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) {
                    const int i = ch * 32 + c16 * 4;
                    *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[i], pk[i + 1], pk[i + 2], pk[i + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> out[token, head*80 ...]
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* orow = p.out + (long long)(row_base + qtok) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_32x16(lane_addr + AT_O_COL + c0, o);
            tmem_ld_wait();
            if (qtok < p.S) {
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                reinterpret_cast<uint4*>(orow + c0)[1] = u1;
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, AT_TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------------------------ 64x64 global, 2 CTAs/SM
// Same algorithm as sam_attn_tcgen05_kernel<64> with 64-key tiles (one image row of keys per tile), sized so that TWO
// CTAs are resident per SM (108 KB shared memory, 256 TMEM columns, <= 128 registers): while one CTA's softmax warps work
// on a score tile the other CTA's MMAs own the tensor pipe, which the single-CTA kernel leaves idle ~60 % of the time.
//   smem  Q 20 KB | 2 x (K 10 KB + V 10 KB) | P 16 KB | rel_h[ky][row] 16 KB | rel_w[kx][row] 16 KB
//         (the rel-pos tables Rh/Rw ride in the two K/V stages during the prologue)
//   TMEM  prologue: T_h [0,128) T_w [128,256)  ->  S double buffer [0,64) [64,128), O [128,208)
constexpr int G2_BK = 64, G2_NS = 2, G2_THREADS = 256;
constexpr int G2_STAGE = 2 * (G2_BK * 128 + G2_BK * 32);          // 20480
constexpr int G2_P = AT_BQ * G2_BK * 2;                           // 16384
constexpr int G2_TAB = 64 * AT_BQ * 2;                            // 16384
constexpr int G2_SMEM = AT_QA + AT_QB + G2_NS * G2_STAGE + G2_P + 2 * G2_TAB + 1024 + 256;
constexpr int G2_O_COL = 128;

__global__ void __launch_bounds__(G2_THREADS, 2)
sam_attn_global64_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                         const __grid_constant__ CUtensorMap tmKVa, const __grid_constant__ CUtensorMap tmKVb,
                         const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                         const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                         const SamAttnParams p) {
    constexpr int KW = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_a = smem;
    uint8_t* q_b = q_a + AT_QA;
    uint8_t* stage0 = q_b + AT_QB;
    auto k_a = [&](int s) { return stage0 + s * G2_STAGE; };
    auto k_b = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 128; };
    auto v_a = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 160; };
    auto v_b = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 288; };
    uint8_t* p_s = stage0 + G2_NS * G2_STAGE;
    bf16* th_s = reinterpret_cast<bf16*>(p_s + G2_P);  // rel_h[ky][row]
    bf16* tw_s = th_s + 64 * AT_BQ;                    // rel_w[kx][row]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tw_s) + G2_TAB);
    uint64_t *q_full = bars, *tab_full = bars + 1, *t_full = bars + 2, *kv_full = bars + 3, *kv_empty = bars + 5,
             *s_full = bars + 7, *s_empty = bars + 9, *p_full = bars + 11, *pv_done = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int E = p.heads * AT_HD;
    const int n_tiles = p.S / G2_BK;
    const int row_base = b * p.S;

    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(tab_full, 1); mbar_init(t_full, 1);
        for (int s = 0; s < G2_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, AT_QA + AT_QB);
            tma_load_2d(q_a, &tmQa, q_full, h * AT_HD, row_base + q0);
            tma_load_2d(q_b, &tmQb, q_full, h * AT_HD + 64, row_base + q0);
            // the two 128-row rel-pos tables (16 KB + 4 KB each) fill the two K/V stages until the prologue MMAs are done
            mbar_arrive_expect_tx(tab_full, 2 * (128 * 128 + 128 * 32));
            tma_load_2d(stage0, &tmRHa, tab_full, 0, 0);
            tma_load_2d(stage0 + 128 * 128, &tmRHb, tab_full, 64, 0);
            tma_load_2d(stage0 + G2_STAGE, &tmRWa, tab_full, 0, 0);
            tma_load_2d(stage0 + G2_STAGE + 128 * 128, &tmRWb, tab_full, 64, 0);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % G2_NS;
                mbar_wait(&kv_empty[s], (j / G2_NS) & 1);
                mbar_arrive_expect_tx(&kv_full[s], G2_STAGE);
                const int r = row_base + j * G2_BK;
                tma_load_2d(k_a(s), &tmKVa, &kv_full[s], E + h * AT_HD, r);
                tma_load_2d(k_b(s), &tmKVb, &kv_full[s], E + h * AT_HD + 64, r);
                tma_load_2d(v_a(s), &tmKVa, &kv_full[s], 2 * E + h * AT_HD, r);
                tma_load_2d(v_b(s), &tmKVb, &kv_full[s], 2 * E + h * AT_HD + 64, r);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t dqa = umma_desc_sw128_kmajor(smem_u32(q_a)), dqb = umma_desc_sw32_kmajor(smem_u32(q_b));
            mbar_wait(q_full, 0);
            mbar_wait(tab_full, 0);
            tc_fence_after();
            issue_qk(tmem_base + 0, q_a, q_b, stage0, stage0 + 128 * 128);                            // T_h
            issue_qk(tmem_base + 128, q_a, q_b, stage0 + G2_STAGE, stage0 + G2_STAGE + 128 * 128);    // T_w
            umma_commit(t_full);
            for (int s = 0; s < G2_NS; ++s) umma_commit(&kv_empty[s]);
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, G2_BK);
            auto issue_s = [&](int j) {
                const int s = j % G2_NS, sb = j & 1;
                mbar_wait(&kv_full[s], (j / G2_NS) & 1);
                mbar_wait(&s_empty[sb], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(k_a(s)));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + sb * G2_BK, dqa + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + sb * G2_BK, dqb, umma_desc_sw32_kmajor(smem_u32(k_b(s))), idesc_s, 1u);
                umma_commit(&s_full[sb]);
            };
            issue_s(0);
            constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                const int s = j % G2_NS;
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < G2_BK / 16; ++ks) {
                    const uint64_t dp = umma_desc_sw128_kmajor(smem_u32(p_s)) + 2 * ks;
                    const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                    umma_bf16(tmem_base + G2_O_COL, dp, umma_desc_sw128_mnmajor(smem_u32(v_a(s) + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + G2_O_COL + 64, dp, umma_desc_sw32_mnmajor(smem_u32(v_b(s) + ks * 512)), idesc16, acc);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
        }
    } else if (warp >= 4) {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const int qtok = q0 + r;
        const int qy = qtok / KW, qx = qtok - qy * KW;
        // ---- prologue: rel_h[ky][r] = T_h[r][qy - ky + 63], rel_w[kx][r] = T_w[r][qx - kx + 63]  (bf16, like the einsums)
        mbar_wait(t_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            bf16* dst = t == 0 ? th_s : tw_s;
            const int qq = (t == 0 ? qy : qx) + KW - 1;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(lane_addr + t * 128 + c0, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int kk = qq - (c0 + i);
                    if (kk >= 0 && kk < KW) dst[kk * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(raw[i]));
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&s_empty[0]); mbar_arrive(&s_empty[1]); }
        const bf16* th_r = th_s + r;
        const bf16* tw_r = tw_s + r;

        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t xr[G2_BK];
#pragma unroll
            for (int c0 = 0; c0 < G2_BK; c0 += 32)
                tmem_ld_32x32(lane_addr + sb * G2_BK + c0, reinterpret_cast<uint32_t(&)[32]>(xr[c0]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            const float rh = __bfloat162float(th_r[j * AT_BQ]);   // key row ky == j
            float x[G2_BK];
            float mx = -INFINITY;
#pragma unroll
            for (int kx = 0; kx < G2_BK; ++kx) {
                const float bias = (__bfloat162float(tw_r[kx * AT_BQ]) + rh) * AT_LOG2E;
                x[kx] = fmaf(__uint_as_float(xr[kx]), p.scale_log2, bias);
                mx = fmaxf(mx, x[kx]);
            }
            float corr = 1.f;
            bool moved = false;
            if (j == 0) {
                m_ref = mx;
            } else if (mx > m_ref + 8.f) {
                corr = ex2_approx(m_ref - mx);
                m_ref = mx;
                moved = true;
            }
            float sum = 0.f;
            uint32_t pk[G2_BK / 2];
#pragma unroll
            for (int c = 0; c < G2_BK; c += 2) {
                const float p0 = ex2_approx(x[c] - m_ref), p1 = ex2_approx(x[c + 1] - m_ref);
                sum += p0 + p1;
                pk[c >> 1] = pack_bf16x2(p0, p1);
            }
            l_run = l_run * corr + sum;
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, moved)) {
#pragma unroll
                    for (int c0 = 0; c0 < AT_HD; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld_32x16(lane_addr + G2_O_COL + c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st_32x16(lane_addr + G2_O_COL + c0, o);
                    }
                    tmem_st_wait();
                }
            }
            uint8_t* row = p_s + r * 128;
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16)
                *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) =
                    make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* orow = p.out + (long long)(row_base + qtok) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_32x16(lane_addr + G2_O_COL + c0, o);
            tmem_ld_wait();
            uint4 u0, u1;
            u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
            u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
            u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
            u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
            reinterpret_cast<uint4*>(orow + c0)[0] = u0;
            reinterpret_cast<uint4*>(orow + c0)[1] = u1;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------ 64x64 global, 2 threads / row
// sam_attn_global64_kernel with EIGHT softmax warps per CTA: every query row is shared by two threads (key columns [0,32) and
// [32,64) of each 64-key tile; the row maximum and the final row sum are exchanged through shared memory, the rel_w values of a
// thread's 32 columns live in 16 registers).  Still two CTAs per SM (setmaxnreg moves registers from the TMA / MMA warps to the
// softmax warps), i.e. four softmax warps per scheduler instead of two.
constexpr int G2H_THREADS = 384;
constexpr int G2H_SMEM = G2_SMEM + 3 * 1024;

__global__ void __launch_bounds__(G2H_THREADS, 2)
sam_attn_global64h_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                         const __grid_constant__ CUtensorMap tmKVa, const __grid_constant__ CUtensorMap tmKVb,
                         const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                         const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                         const SamAttnParams p) {
    constexpr int KW = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_a = smem;
    uint8_t* q_b = q_a + AT_QA;
    uint8_t* stage0 = q_b + AT_QB;
    auto k_a = [&](int s) { return stage0 + s * G2_STAGE; };
    auto k_b = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 128; };
    auto v_a = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 160; };
    auto v_b = [&](int s) { return stage0 + s * G2_STAGE + G2_BK * 288; };
    uint8_t* p_s = stage0 + G2_NS * G2_STAGE;
    bf16* th_s = reinterpret_cast<bf16*>(p_s + G2_P);  // rel_h[ky][row]
    bf16* tw_s = th_s + 64 * AT_BQ;                    // rel_w[kx][row]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tw_s) + G2_TAB);
    // K and V of a stage have their own barriers: a K slot is free as soon as S = Q K^T of its tile has completed (two tiles
    // before the slot is needed again), a V slot only after P V -- with one barrier pair per stage the next K load could not
    // start before the previous tile's P V had finished and its latency sat on the critical path of every tile
    uint64_t *q_full = bars, *tab_full = bars + 1, *t_full = bars + 2, *k_full = bars + 3, *k_empty = bars + 5,
             *s_full = bars + 7, *s_empty = bars + 9, *p_full = bars + 11, *pv_done = bars + 12, *v_full = bars + 13,
             *v_empty = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    float* mxbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [tile parity][half][row]
    float* lbuf = mxbuf + 4 * AT_BQ;                                                   // [half][row]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int E = p.heads * AT_HD;
    const int n_tiles = p.S / G2_BK;
    const int row_base = b * p.S;

    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(tab_full, 1); mbar_init(t_full, 1);
        for (int s = 0; s < G2_NS; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 8); }
        mbar_init(p_full, 8);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // warps 0-3 (TMA, MMA issue, TMEM allocation) hand registers to the eight softmax warps
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, AT_QA + AT_QB);
            tma_load_2d(q_a, &tmQa, q_full, h * AT_HD, row_base + q0);
            tma_load_2d(q_b, &tmQb, q_full, h * AT_HD + 64, row_base + q0);
            // the two 128-row rel-pos tables (16 KB + 4 KB each) fill the two K/V stages until the prologue MMAs are done
            mbar_arrive_expect_tx(tab_full, 2 * (128 * 128 + 128 * 32));
            tma_load_2d(stage0, &tmRHa, tab_full, 0, 0);
            tma_load_2d(stage0 + 128 * 128, &tmRHb, tab_full, 64, 0);
            tma_load_2d(stage0 + G2_STAGE, &tmRWa, tab_full, 0, 0);
            tma_load_2d(stage0 + G2_STAGE + 128 * 128, &tmRWb, tab_full, 64, 0);
            for (int j = 0; j < n_tiles; ++j) {   // K ring (warp 3 feeds the V ring)
                const int s = j % G2_NS;
                mbar_wait(&k_empty[s], (j / G2_NS) & 1);
                mbar_arrive_expect_tx(&k_full[s], G2_STAGE / 2);
                const int r = row_base + j * G2_BK;
                tma_load_2d(k_a(s), &tmKVa, &k_full[s], E + h * AT_HD, r);
                tma_load_2d(k_b(s), &tmKVb, &k_full[s], E + h * AT_HD + 64, r);
            }
        }
    } else if (warp == 3) {
        if (lane == 0) {
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % G2_NS;
                mbar_wait(&v_empty[s], (j / G2_NS) & 1);
                mbar_arrive_expect_tx(&v_full[s], G2_STAGE / 2);
                const int r = row_base + j * G2_BK;
                tma_load_2d(v_a(s), &tmKVa, &v_full[s], 2 * E + h * AT_HD, r);
                tma_load_2d(v_b(s), &tmKVb, &v_full[s], 2 * E + h * AT_HD + 64, r);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t dqa = umma_desc_sw128_kmajor(smem_u32(q_a)), dqb = umma_desc_sw32_kmajor(smem_u32(q_b));
            mbar_wait(q_full, 0);
            mbar_wait(tab_full, 0);
            tc_fence_after();
            issue_qk(tmem_base + 0, q_a, q_b, stage0, stage0 + 128 * 128);                            // T_h
            issue_qk(tmem_base + 128, q_a, q_b, stage0 + G2_STAGE, stage0 + G2_STAGE + 128 * 128);    // T_w
            umma_commit(t_full);
            for (int s = 0; s < G2_NS; ++s) { umma_commit(&k_empty[s]); umma_commit(&v_empty[s]); }
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, G2_BK);
            auto issue_s = [&](int j) {
                const int s = j % G2_NS, sb = j & 1;
                mbar_wait(&k_full[s], (j / G2_NS) & 1);
                mbar_wait(&s_empty[sb], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(k_a(s)));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + sb * G2_BK, dqa + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + sb * G2_BK, dqb, umma_desc_sw32_kmajor(smem_u32(k_b(s))), idesc_s, 1u);
                umma_commit(&s_full[sb]);
                umma_commit(&k_empty[s]);
            };
            issue_s(0);
            constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                const int s = j % G2_NS;
                mbar_wait(&v_full[s], (j / G2_NS) & 1);
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < G2_BK / 16; ++ks) {
                    const uint64_t dp = umma_desc_sw128_kmajor(smem_u32(p_s)) + 2 * ks;
                    const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                    umma_bf16(tmem_base + G2_O_COL, dp, umma_desc_sw128_mnmajor(smem_u32(v_a(s) + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + G2_O_COL + 64, dp, umma_desc_sw32_mnmajor(smem_u32(v_b(s) + ks * 512)), idesc16, acc);
                }
                umma_commit(&v_empty[s]);
                umma_commit(pv_done);
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        // two threads per query row: warps 4-7 (half 0) take key columns [0,32) of every 64-key tile, warps 8-11 (half 1) take
        // [32,64); warp w and w+4 share a TMEM lane quadrant.  Twice the warps per scheduler hide the TMEM / barrier latencies
        // that left the issue slots half empty with one thread per row.
        const int quad = warp & 3, half = (warp >> 2) - 1;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const int qtok = q0 + r;
        const int qy = qtok / KW, qx = qtok - qy * KW;
        // ---- prologue: half 0 files rel_h[ky][r] = T_h[r][qy - ky + 63], half 1 files rel_w[kx][r] = T_w[r][qx - kx + 63]
        mbar_wait(t_full, 0);
        tc_fence_after();
        {
            bf16* dst = half == 0 ? th_s : tw_s;
            const int qq = (half == 0 ? qy : qx) + KW - 1;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(lane_addr + half * 128 + c0, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int kk = qq - (c0 + i);
                    if (kk >= 0 && kk < KW) dst[kk * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(raw[i]));
                }
            }
        }
        tc_fence_before();
        named_bar_sync(1, 256);   // both tables complete (each half reads what the other wrote); T_h / T_w columns are free
        if (lane == 0) { mbar_arrive(&s_empty[0]); mbar_arrive(&s_empty[1]); }
        const bf16* th_r = th_s + r;
        // this thread's 32 rel_w values (the same for every key row): packed bf16 pairs in 16 registers
        uint32_t twp[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t lo = *reinterpret_cast<const uint16_t*>(tw_s + (half * 32 + 2 * i) * AT_BQ + r);
            const uint32_t hi = *reinterpret_cast<const uint16_t*>(tw_s + (half * 32 + 2 * i + 1) * AT_BQ + r);
            twp[i] = lo | (hi << 16);
        }

        // Single pass per tile: the scores are formed directly relative to the running reference m_ref (log2 domain), exponentiated
        // at once, and the reference only moves when a tile's maximum exceeds it by more than 2^8 (then the tile's probabilities,
        // the row sum and O are scaled by 2^-excess -- rare after the first tiles).  About 6.5 instructions per score element
        // instead of 12: unpack rel_w, two FMAs, max, ex2, sum, half a pack.
        float m_ref = 0.f, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t xr[32];
            tmem_ld_32x32(lane_addr + sb * G2_BK + half * 32, xr);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            const float rh = __bfloat162float(th_r[j * AT_BQ]);   // key row ky == j
            const float cbase = fmaf(rh, AT_LOG2E, -m_ref);
            float mx = -INFINITY;
            // two score elements per issued instruction (FFMA2): the loop is issue-bound, the arithmetic per element is unchanged
            const uint64_t cbase2 = f32x2_pack(cbase, cbase), l2e2 = f32x2_pack(AT_LOG2E, AT_LOG2E), sc2 = f32x2_pack(p.scale_log2, p.scale_log2);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const uint64_t tw2 = f32x2_pack(__uint_as_float(twp[i >> 1] << 16), __uint_as_float(twp[i >> 1] & 0xffff0000u));
                const uint64_t s2 = f32x2_pack(__uint_as_float(xr[i]), __uint_as_float(xr[i + 1]));
                float x0, x1;
                f32x2_unpack(f32x2_fma(s2, sc2, f32x2_fma(tw2, l2e2, cbase2)), x0, x1);
                mx = fmaxf(mx, fmaxf(x0, x1));
                xr[i] = __float_as_uint(ex2_approx(x0));
                xr[i + 1] = __float_as_uint(ex2_approx(x1));
            }
            // row maximum over both halves (relative to m_ref)
            float* mb = mxbuf + (j & 1) * 2 * AT_BQ;
            mb[half * AT_BQ + r] = mx;
            named_bar_sync(2 + quad, 64);
            mx = fmaxf(mx, mb[(half ^ 1) * AT_BQ + r]);
            float corr = 1.f;
            const bool moved = (j == 0) || (mx > 8.f);
            if (moved) {
                corr = ex2_approx(-mx);
                m_ref += mx;
                const uint64_t corr2 = f32x2_pack(corr, corr);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float a0, a1;
                    f32x2_unpack(f32x2_mul(f32x2_pack(__uint_as_float(xr[i]), __uint_as_float(xr[i + 1])), corr2), a0, a1);
                    xr[i] = __float_as_uint(a0);
                    xr[i + 1] = __float_as_uint(a1);
                }
            }
            uint64_t sum2 = f32x2_pack(0.f, 0.f);
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float p0 = __uint_as_float(xr[c]), p1 = __uint_as_float(xr[c + 1]);
                sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                pk[c >> 1] = pack_bf16x2(p0, p1);
            }
            float sum_lo, sum_hi;
            f32x2_unpack(sum2, sum_lo, sum_hi);
            const float sum = sum_lo + sum_hi;
            l_run = l_run * corr + sum;
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, moved)) {   // the partner warp sees the same rows and takes the same branch
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const int c0 = half * 48 + cc * 16;
                        if (c0 < AT_HD) {
                            uint32_t o[16];
                            tmem_ld_32x16(lane_addr + G2_O_COL + c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                            tmem_st_32x16(lane_addr + G2_O_COL + c0, o);
                        }
                    }
                    tmem_st_wait();
                }
            }
            uint8_t* row = p_s + r * 128;
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16)
                *reinterpret_cast<uint4*>(row + (((half * 4 + c16) ^ (r & 7)) << 4)) =
                    make_uint4(pk[c16 * 4], pk[c16 * 4 + 1], pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        lbuf[half * AT_BQ + r] = l_run;
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        named_bar_sync(2 + quad, 64);
        const float l_tot = lbuf[r] + lbuf[AT_BQ + r];
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        bf16* orow = p.out + (long long)(row_base + qtok) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int c0 = half * 48 + cc * 16;
            if (c0 < AT_HD) {
                uint32_t o[16];
                tmem_ld_32x16(lane_addr + G2_O_COL + c0, o);
                tmem_ld_wait();
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                reinterpret_cast<uint4*>(orow + c0)[1] = u1;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------ 14x14 windows
// Window attention has 196 keys: ONE key tile (UMMA N = 208), no online softmax, and so little work per (window, head,
// query tile) that latency dominates.  This variant is sized for TWO resident CTAs per SM (112 KB shared memory, 256
// TMEM columns, 160 threads) so that one CTA's TMA / MMA / barrier latencies hide under the other's softmax:
//   TMEM  [0,32) T_h, [32,64) T_w  ->  [0,208) S (bias-added scores are written back in place)  ->  [0,80) O
//   smem  Q (20 KB) + K (33 KB) are dead once S is committed: P (52 KB, three 64-key chunks + one 16-key chunk) aliases them.
constexpr int WN_KEYS = 208, WN_S = 196, WN_KW = 14, WN_THREADS = 160;
constexpr int WN_KA = WN_KEYS * 128, WN_KB = WN_KEYS * 32;                 // 26624, 6656
constexpr int WN_OFF_QA = 0, WN_OFF_QB = 16384, WN_OFF_KA = 20480, WN_OFF_KB = WN_OFF_KA + WN_KA;       // 47104
constexpr int WN_OFF_VA = 54272, WN_OFF_VB = WN_OFF_VA + WN_KA;                                         // 80896
constexpr int WN_OFF_RHA = 88064, WN_OFF_RHB = WN_OFF_RHA + 4096, WN_OFF_RWA = 93184, WN_OFF_RWB = WN_OFF_RWA + 4096;
constexpr int WN_OFF_TH = 98304, WN_OFF_TW = WN_OFF_TH + 28 * 256, WN_OFF_BAR = WN_OFF_TW + 28 * 256;  // 112640
constexpr int WN_OFF_P3 = 49152;                                           // 16-key P chunk (32B-swizzled)
constexpr int WN_SMEM = WN_OFF_BAR + 256 + 1024;
static_assert(WN_OFF_KB + WN_KB <= WN_OFF_VA && WN_OFF_VB + WN_KB <= WN_OFF_RHA && WN_OFF_P3 + 4096 <= WN_OFF_VA, "smem map");

template <int C0, int N>
IVLM_DEVINL float win_scores(uint32_t (&raw)[N], const float (&rh)[WN_KW], const float (&rw)[WN_KW], float scale_log2, float mx) {
    // two keys per FFMA2 (N and the 196-key limit are even: a pair is valid or masked as a whole)
    const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        const int k = C0 + i;  // compile-time after unrolling
        float x0, x1;
        if (k < WN_S) {
            const uint64_t bias2 = f32x2_pack(rh[k / WN_KW] + rw[k % WN_KW], rh[(k + 1) / WN_KW] + rw[(k + 1) % WN_KW]);
            f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), sc2, bias2), x0, x1);
            mx = fmaxf(mx, fmaxf(x0, x1));
        } else {
            x0 = x1 = -INFINITY;
        }
        raw[i] = __float_as_uint(x0);
        raw[i + 1] = __float_as_uint(x1);
    }
    return mx;
}

__global__ void __launch_bounds__(WN_THREADS, 2)
sam_attn_window_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                               const __grid_constant__ CUtensorMap tmKVa, const __grid_constant__ CUtensorMap tmKVb,
                               const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                               const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                               const SamAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    bf16* th_s = reinterpret_cast<bf16*>(smem + WN_OFF_TH);  // [28 idx][128 rows]
    bf16* tw_s = reinterpret_cast<bf16*>(smem + WN_OFF_TW);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WN_OFF_BAR);
    uint64_t *qt_full = bars, *k_full = bars + 1, *v_full = bars + 2, *t_full = bars + 3, *t_read = bars + 4,
             *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int E = p.heads * AT_HD;
    const int row_base = b * WN_S;

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(qt_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(t_full, 1);
            mbar_init(t_read, 4); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ control: TMA + MMA issue
        if (lane == 0) {
            mbar_arrive_expect_tx(qt_full, 20480 + 2 * (4096 + 1024));
            tma_load_2d(smem + WN_OFF_QA, &tmQa, qt_full, h * AT_HD, row_base + q0);
            tma_load_2d(smem + WN_OFF_QB, &tmQb, qt_full, h * AT_HD + 64, row_base + q0);
            tma_load_2d(smem + WN_OFF_RHA, &tmRHa, qt_full, 0, 0);
            tma_load_2d(smem + WN_OFF_RHB, &tmRHb, qt_full, 64, 0);
            tma_load_2d(smem + WN_OFF_RWA, &tmRWa, qt_full, 0, 0);
            tma_load_2d(smem + WN_OFF_RWB, &tmRWb, qt_full, 64, 0);
            mbar_arrive_expect_tx(k_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_KA, &tmKVa, k_full, E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_KB, &tmKVb, k_full, E + h * AT_HD + 64, row_base);
            mbar_arrive_expect_tx(v_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_VA, &tmKVa, v_full, 2 * E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_VB, &tmKVb, v_full, 2 * E + h * AT_HD + 64, row_base);
            // Optional (option "attn_prefetch_ahead", off by default): pull the q / k / v tiles of a CTA further down the launch
            // order into L2.  The kernel's top stalls are barrier waits behind its TMA loads and MMAs, but the experiment was
            // neutral -- the latency that paces a CTA is the serial chain load -> T -> S -> softmax -> P V, not the DRAM access.
            if (p.prefetch_ahead > 0) {
                const long long lin = (long long)blockIdx.x + 2LL * (h + (long long)p.heads * b) + p.prefetch_ahead;
                const int nb = (int)(lin / (2LL * p.heads)), nh = (int)((lin >> 1) % p.heads), nx = (int)(lin & 1);
                if (nb < (int)gridDim.z) {
                    const int nrow = nb * WN_S;
                    tma_prefetch_2d_l2(&tmQa, nh * AT_HD, nrow + nx * AT_BQ);
                    tma_prefetch_2d_l2(&tmQb, nh * AT_HD + 64, nrow + nx * AT_BQ);
                    if (nx == 0) {   // both query tiles of a (window, head) read the same keys / values
                        tma_prefetch_2d_l2(&tmKVa, E + nh * AT_HD, nrow);
                        tma_prefetch_2d_l2(&tmKVb, E + nh * AT_HD + 64, nrow);
                        tma_prefetch_2d_l2(&tmKVa, 2 * E + nh * AT_HD, nrow);
                        tma_prefetch_2d_l2(&tmKVb, 2 * E + nh * AT_HD + 64, nrow);
                    }
                }
            }

            const uint64_t dqa = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_QA));
            const uint64_t dqb = umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_QB));
            mbar_wait(qt_full, 0);
            tc_fence_after();
            {   // T_h -> cols [0,32), T_w -> cols [32,64): Q . table^T, N = 32 (27 table rows + zero fill)
                constexpr uint32_t idesc = umma_idesc_bf16(128, 32);
                const uint64_t dh = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RHA));
                const uint64_t dw = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RWA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dqa + 2 * k, dh + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RHB)), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 32, dqa + 2 * k, dw + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + 32, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RWB)), idesc, 1u);
                umma_commit(t_full);
            }
            mbar_wait(k_full, 0);
            mbar_wait(t_read, 0);  // the softmax warps have taken T_h / T_w out of the columns S overwrites
            tc_fence_after();
            {   // S = Q K^T, N = 208
                constexpr uint32_t idesc = umma_idesc_bf16(128, WN_KEYS);
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_KA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dqa + 2 * k, dk + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_KB)), idesc, 1u);
                umma_commit(s_full);
            }
            mbar_wait(v_full, 0);
            mbar_wait(p_full, 0);
            tc_fence_after();
            {   // O = P V over 13 key steps of 16
                constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
#pragma unroll
                for (int ks = 0; ks < WN_KEYS / 16; ++ks) {
                    const uint64_t dp = ks < 12 ? umma_desc_sw128_kmajor(smem_u32(smem + (ks >> 2) * 16384)) + 2 * (ks & 3)
                                                : umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_P3));
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    umma_bf16(tmem_base, dp, umma_desc_sw128_mnmajor(smem_u32(smem + WN_OFF_VA + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + 64, dp, umma_desc_sw32_mnmajor(smem_u32(smem + WN_OFF_VB + ks * 512)), idesc16, acc);
                }
                umma_commit(o_full);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps (quadrant == warp)
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
        const int qtok = q0 + r;
        const bool q_ok = qtok < WN_S;
        const int qy = q_ok ? qtok / WN_KW : 0, qx = q_ok ? qtok - (qtok / WN_KW) * WN_KW : 0;
        mbar_wait(t_full, 0);
        tc_fence_after();
        {
            uint32_t th[32], tw[32];
            tmem_ld_32x32(lane_addr, th);
            tmem_ld_32x32(lane_addr + 32, tw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 28; ++i) {
                th_s[i * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(th[i]));
                tw_s[i * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(tw[i]));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_read);
        float rh[WN_KW], rw[WN_KW];  // this row's rel_h[ky], rel_w[kx] (bf16 values), pre-multiplied by log2(e)
#pragma unroll
        for (int i = 0; i < WN_KW; ++i) {
            rh[i] = __bfloat162float(th_s[(qy - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
            rw[i] = __bfloat162float(tw_s[(qx - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
        }
        mbar_wait(s_full, 0);
        tc_fence_after();
        // ---- pass 1: x = s*scale + bias (log2 domain) written back in place, row max.  Both passes are software
        // pipelined over two register buffers: the TMEM read of chunk c+1 is issued before chunk c is processed, so its
        // latency hides under the arithmetic (two softmax warps per scheduler do not hide it otherwise).
        float mx = -INFINITY;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(lane_addr, ra);
#define WIN_PASS1(C0, CUR, NXT, NEXT_LD)                                            \
        {                                                                           \
            tmem_ld_wait();                                                         \
            NEXT_LD;                                                                \
            mx = win_scores<C0, 32>(CUR, rh, rw, p.scale_log2, mx);                 \
            tmem_st_32x32(lane_addr + C0, CUR);                                     \
        }
        WIN_PASS1(0, ra, rb, tmem_ld_32x32(lane_addr + 32, rb))
        WIN_PASS1(32, rb, ra, tmem_ld_32x32(lane_addr + 64, ra))
        WIN_PASS1(64, ra, rb, tmem_ld_32x32(lane_addr + 96, rb))
        WIN_PASS1(96, rb, ra, tmem_ld_32x32(lane_addr + 128, ra))
        WIN_PASS1(128, ra, rb, tmem_ld_32x32(lane_addr + 160, rb))
        uint32_t rc[16];
        WIN_PASS1(160, rb, ra, tmem_ld_32x16(lane_addr + 192, rc))
#undef WIN_PASS1
        tmem_ld_wait();
        mx = win_scores<192, 16>(rc, rh, rw, p.scale_log2, mx);
        tmem_st_32x16(lane_addr + 192, rc);
        tmem_st_wait();
        // ---- pass 2: p = 2^(x - max) -> bf16 -> swizzled P (aliases the dead Q / K tiles)
        uint64_t sum2 = f32x2_pack(0.f, 0.f);            // two partial row sums (FADD2), combined after the pass
        const uint64_t nmx2 = f32x2_pack(-mx, -mx);
        auto exp_chunk = [&](const uint32_t(&raw)[32], int c0) {
            uint8_t* row = smem + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float d0, d1;
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(raw[j8 * 8 + 2 * e]), __uint_as_float(raw[j8 * 8 + 2 * e + 1])), nmx2), d0, d1);
                    const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
                    sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
                const int c16 = ((c0 & 63) >> 3) + j8;
                *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        };
        tmem_ld_32x32(lane_addr, ra);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 32, rb);  exp_chunk(ra, 0);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 64, ra);  exp_chunk(rb, 32);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 96, rb);  exp_chunk(ra, 64);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 128, ra); exp_chunk(rb, 96);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 160, rb); exp_chunk(ra, 128);
        tmem_ld_wait(); tmem_ld_32x16(lane_addr + 192, rc); exp_chunk(rb, 160);
        {
            tmem_ld_wait();
            uint8_t* row = smem + WN_OFF_P3 + r * 32;
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float d0, d1;
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(rc[j8 * 8 + 2 * e]), __uint_as_float(rc[j8 * 8 + 2 * e + 1])), nmx2), d0, d1);
                    const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
                    sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
                *reinterpret_cast<uint4*>(row + ((j8 ^ ((r >> 2) & 1)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue
        mbar_wait(o_full, 0);
        tc_fence_after();
        float sum_lo, sum_hi;
        f32x2_unpack(sum2, sum_lo, sum_hi);
        const float sum = sum_lo + sum_hi;
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        // window_unpartition (image_encoder.py:291-318) folded into the store: with out_map the row goes straight to its token
        // position and the rows of the zero padding are never written
        long long orow_i = q_ok ? (long long)(row_base + qtok) : -1;
        if (q_ok && p.out_map != nullptr) orow_i = p.out_map[row_base + qtok];
        const bool st_ok = orow_i >= 0;
        bf16* orow = p.out + (st_ok ? orow_i : 0) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_32x16(lane_addr + c0, o);
            tmem_ld_wait();
            if (st_ok) {
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                reinterpret_cast<uint4*>(orow + c0)[1] = u1;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------ 14x14 windows, persistent
// sam_attn_window_tcgen05_kernel as a persistent kernel: 2 CTAs per SM loop over the (query tile, head, window) items.  Per item
// the chain is still load -> T -> S -> two softmax passes -> P V -> store, but (a) barrier / TMEM set-up and the rel-pos tables
// are paid once per CTA, (b) the next item's Q / K / V loads and its T product (parked in TMEM columns [192,256), clear of the
// previous O) run under the previous item's epilogue.  Same arithmetic in the same order: bit-identical outputs.
__global__ void __launch_bounds__(WN_THREADS, 2)
sam_attn_window_persist_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                               const __grid_constant__ CUtensorMap tmKVa, const __grid_constant__ CUtensorMap tmKVb,
                               const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                               const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                               const SamAttnParams p, const int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    bf16* th_s = reinterpret_cast<bf16*>(smem + WN_OFF_TH);  // [28 idx][128 rows]
    bf16* tw_s = reinterpret_cast<bf16*>(smem + WN_OFF_TW);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WN_OFF_BAR);
    uint64_t *qt_full = bars, *k_full = bars + 1, *v_full = bars + 2, *t_full = bars + 3, *t_read = bars + 4,
             *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7;
    uint64_t* o_read = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int E = p.heads * AT_HD;

    if (warp == 4) {
        if (lane == 0) {
            mbar_init(qt_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(t_full, 1);
            mbar_init(t_read, 4); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_read, 4);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ control: TMA + MMA issue
        if (lane == 0) {
          for (int it = 0, item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            const int q0 = (item & 1) * AT_BQ, h = (item >> 1) % p.heads, b = (item >> 1) / p.heads;
            const int row_base = b * WN_S;
            // Q / K (aliased by P) and V of the previous item are free once its P.V has retired; its epilogue (TMEM -> global) is
            // still running on the softmax warps while these loads are in flight
            if (it > 0) mbar_wait(o_full, ph ^ 1);
            mbar_arrive_expect_tx(qt_full, it == 0 ? 20480 + 2 * (4096 + 1024) : 20480);
            tma_load_2d(smem + WN_OFF_QA, &tmQa, qt_full, h * AT_HD, row_base + q0);
            tma_load_2d(smem + WN_OFF_QB, &tmQb, qt_full, h * AT_HD + 64, row_base + q0);
            if (it == 0) {   // the rel-pos tables are the same for every item: loaded once per CTA
                tma_load_2d(smem + WN_OFF_RHA, &tmRHa, qt_full, 0, 0);
                tma_load_2d(smem + WN_OFF_RHB, &tmRHb, qt_full, 64, 0);
                tma_load_2d(smem + WN_OFF_RWA, &tmRWa, qt_full, 0, 0);
                tma_load_2d(smem + WN_OFF_RWB, &tmRWb, qt_full, 64, 0);
            }
            mbar_arrive_expect_tx(k_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_KA, &tmKVa, k_full, E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_KB, &tmKVb, k_full, E + h * AT_HD + 64, row_base);
            mbar_arrive_expect_tx(v_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_VA, &tmKVa, v_full, 2 * E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_VB, &tmKVb, v_full, 2 * E + h * AT_HD + 64, row_base);
            const uint64_t dqa = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_QA));
            const uint64_t dqb = umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_QB));
            mbar_wait(qt_full, ph);
            tc_fence_after();
            {   // T_h -> cols [192,224), T_w -> cols [224,256): clear of the previous item's O in [0,80), which its epilogue still reads
                constexpr uint32_t idesc = umma_idesc_bf16(128, 32);
                const uint64_t dh = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RHA));
                const uint64_t dw = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RWA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 192, dqa + 2 * k, dh + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + 192, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RHB)), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 224, dqa + 2 * k, dw + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + 224, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RWB)), idesc, 1u);
                umma_commit(t_full);
            }
            mbar_wait(k_full, ph);
            mbar_wait(t_read, ph);  // the softmax warps have taken T_h / T_w out of the columns S overwrites
            if (it > 0) mbar_wait(o_read, ph ^ 1);   // ... and the previous item's O out of [0,80)
            tc_fence_after();
            {   // S = Q K^T, N = 208
                constexpr uint32_t idesc = umma_idesc_bf16(128, WN_KEYS);
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_KA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dqa + 2 * k, dk + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_KB)), idesc, 1u);
                umma_commit(s_full);
            }
            mbar_wait(v_full, ph);
            mbar_wait(p_full, ph);
            tc_fence_after();
            {   // O = P V over 13 key steps of 16
                constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
#pragma unroll
                for (int ks = 0; ks < WN_KEYS / 16; ++ks) {
                    const uint64_t dp = ks < 12 ? umma_desc_sw128_kmajor(smem_u32(smem + (ks >> 2) * 16384)) + 2 * (ks & 3)
                                                : umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_P3));
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    umma_bf16(tmem_base, dp, umma_desc_sw128_mnmajor(smem_u32(smem + WN_OFF_VA + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + 64, dp, umma_desc_sw32_mnmajor(smem_u32(smem + WN_OFF_VB + ks * 512)), idesc16, acc);
                }
                umma_commit(o_full);
            }
          }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps (quadrant == warp)
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
      for (int it = 0, item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int q0 = (item & 1) * AT_BQ, h = (item >> 1) % p.heads, b = (item >> 1) / p.heads;
        const int row_base = b * WN_S;
        const int qtok = q0 + r;
        const bool q_ok = qtok < WN_S;
        const int qy = q_ok ? qtok / WN_KW : 0, qx = q_ok ? qtok - (qtok / WN_KW) * WN_KW : 0;
        mbar_wait(t_full, ph);
        tc_fence_after();
        {
            uint32_t th[32], tw[32];
            tmem_ld_32x32(lane_addr + 192, th);
            tmem_ld_32x32(lane_addr + 224, tw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 28; ++i) {
                th_s[i * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(th[i]));
                tw_s[i * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(tw[i]));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_read);
        float rh[WN_KW], rw[WN_KW];  // this row's rel_h[ky], rel_w[kx] (bf16 values), pre-multiplied by log2(e)
#pragma unroll
        for (int i = 0; i < WN_KW; ++i) {
            rh[i] = __bfloat162float(th_s[(qy - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
            rw[i] = __bfloat162float(tw_s[(qx - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
        }
        mbar_wait(s_full, ph);
        tc_fence_after();
        // ---- pass 1: x = s*scale + bias (log2 domain) written back in place, row max.  Both passes are software
        // pipelined over two register buffers: the TMEM read of chunk c+1 is issued before chunk c is processed, so its
        // latency hides under the arithmetic (two softmax warps per scheduler do not hide it otherwise).
        float mx = -INFINITY;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(lane_addr, ra);
#define WIN_PASS1(C0, CUR, NXT, NEXT_LD)                                            \
        {                                                                           \
            tmem_ld_wait();                                                         \
            NEXT_LD;                                                                \
            mx = win_scores<C0, 32>(CUR, rh, rw, p.scale_log2, mx);                 \
            tmem_st_32x32(lane_addr + C0, CUR);                                     \
        }
        WIN_PASS1(0, ra, rb, tmem_ld_32x32(lane_addr + 32, rb))
        WIN_PASS1(32, rb, ra, tmem_ld_32x32(lane_addr + 64, ra))
        WIN_PASS1(64, ra, rb, tmem_ld_32x32(lane_addr + 96, rb))
        WIN_PASS1(96, rb, ra, tmem_ld_32x32(lane_addr + 128, ra))
        WIN_PASS1(128, ra, rb, tmem_ld_32x32(lane_addr + 160, rb))
        uint32_t rc[16];
        WIN_PASS1(160, rb, ra, tmem_ld_32x16(lane_addr + 192, rc))
#undef WIN_PASS1
        tmem_ld_wait();
        mx = win_scores<192, 16>(rc, rh, rw, p.scale_log2, mx);
        tmem_st_32x16(lane_addr + 192, rc);
        tmem_st_wait();
        // ---- pass 2: p = 2^(x - max) -> bf16 -> swizzled P (aliases the dead Q / K tiles)
        uint64_t sum2 = f32x2_pack(0.f, 0.f);            // two partial row sums (FADD2), combined after the pass
        const uint64_t nmx2 = f32x2_pack(-mx, -mx);
        auto exp_chunk = [&](const uint32_t(&raw)[32], int c0) {
            uint8_t* row = smem + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float d0, d1;
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(raw[j8 * 8 + 2 * e]), __uint_as_float(raw[j8 * 8 + 2 * e + 1])), nmx2), d0, d1);
                    const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
                    sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
                const int c16 = ((c0 & 63) >> 3) + j8;
                *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        };
        tmem_ld_32x32(lane_addr, ra);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 32, rb);  exp_chunk(ra, 0);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 64, ra);  exp_chunk(rb, 32);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 96, rb);  exp_chunk(ra, 64);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 128, ra); exp_chunk(rb, 96);
        tmem_ld_wait(); tmem_ld_32x32(lane_addr + 160, rb); exp_chunk(ra, 128);
        tmem_ld_wait(); tmem_ld_32x16(lane_addr + 192, rc); exp_chunk(rb, 160);
        {
            tmem_ld_wait();
            uint8_t* row = smem + WN_OFF_P3 + r * 32;
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float d0, d1;
                    f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(rc[j8 * 8 + 2 * e]), __uint_as_float(rc[j8 * 8 + 2 * e + 1])), nmx2), d0, d1);
                    const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
                    sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                    pk[e] = pack_bf16x2(p0, p1);
                }
                *reinterpret_cast<uint4*>(row + ((j8 ^ ((r >> 2) & 1)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue
        mbar_wait(o_full, ph);
        tc_fence_after();
        float sum_lo, sum_hi;
        f32x2_unpack(sum2, sum_lo, sum_hi);
        const float sum = sum_lo + sum_hi;
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        // window_unpartition (image_encoder.py:291-318) folded into the store: with out_map the row goes straight to its token
        // position and the rows of the zero padding are never written
        long long orow_i = q_ok ? (long long)(row_base + qtok) : -1;
        if (q_ok && p.out_map != nullptr) orow_i = p.out_map[row_base + qtok];
        const bool st_ok = orow_i >= 0;
        bf16* orow = p.out + (st_ok ? orow_i : 0) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int c0 = 0; c0 < AT_HD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_32x16(lane_addr + c0, o);
            tmem_ld_wait();
            if (st_ok) {
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                reinterpret_cast<uint4*>(orow + c0)[1] = u1;
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_read);   // this warp's reads of O are done: the next item's S may overwrite the columns
      }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}


// ------------------------------------------------------------------------------------------------ 14x14 windows, 2 threads / row
// sam_attn_window_tcgen05_kernel with eight softmax warps (two threads per query row: key columns [0,96) and [96,208)); the
// row maximum and row sum are exchanged through shared memory.  Still two CTAs per SM.
constexpr int WNH_THREADS = 288;
constexpr int WNH_SMEM = WN_SMEM;   // the exchange buffer aliases a dead tile: two CTAs still fit per SM

__global__ void __launch_bounds__(WNH_THREADS, 2)
sam_attn_window_h_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb,
                               const __grid_constant__ CUtensorMap tmKVa, const __grid_constant__ CUtensorMap tmKVb,
                               const __grid_constant__ CUtensorMap tmRHa, const __grid_constant__ CUtensorMap tmRHb,
                               const __grid_constant__ CUtensorMap tmRWa, const __grid_constant__ CUtensorMap tmRWb,
                               const SamAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    bf16* th_s = reinterpret_cast<bf16*>(smem + WN_OFF_TH);  // [28 idx][128 rows]
    bf16* tw_s = reinterpret_cast<bf16*>(smem + WN_OFF_TW);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WN_OFF_BAR);
    uint64_t *qt_full = bars, *k_full = bars + 1, *v_full = bars + 2, *t_full = bars + 3, *t_read = bars + 4,
             *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* xbuf = reinterpret_cast<float*>(smem + WN_OFF_RHA);   // [max | sum][half][row]: over the rel-pos table tile, dead once T is done

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int E = p.heads * AT_HD;
    const int row_base = b * WN_S;

    if (warp == 8) {
        if (lane == 0) {
            mbar_init(qt_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(t_full, 1);
            mbar_init(t_read, 8); mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(o_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ------------------------------------------------------------------ control: TMA + MMA issue
        if (lane == 0) {
            mbar_arrive_expect_tx(qt_full, 20480 + 2 * (4096 + 1024));
            tma_load_2d(smem + WN_OFF_QA, &tmQa, qt_full, h * AT_HD, row_base + q0);
            tma_load_2d(smem + WN_OFF_QB, &tmQb, qt_full, h * AT_HD + 64, row_base + q0);
            tma_load_2d(smem + WN_OFF_RHA, &tmRHa, qt_full, 0, 0);
            tma_load_2d(smem + WN_OFF_RHB, &tmRHb, qt_full, 64, 0);
            tma_load_2d(smem + WN_OFF_RWA, &tmRWa, qt_full, 0, 0);
            tma_load_2d(smem + WN_OFF_RWB, &tmRWb, qt_full, 64, 0);
            mbar_arrive_expect_tx(k_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_KA, &tmKVa, k_full, E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_KB, &tmKVb, k_full, E + h * AT_HD + 64, row_base);
            mbar_arrive_expect_tx(v_full, WN_KA + WN_KB);
            tma_load_2d(smem + WN_OFF_VA, &tmKVa, v_full, 2 * E + h * AT_HD, row_base);
            tma_load_2d(smem + WN_OFF_VB, &tmKVb, v_full, 2 * E + h * AT_HD + 64, row_base);

            const uint64_t dqa = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_QA));
            const uint64_t dqb = umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_QB));
            mbar_wait(qt_full, 0);
            tc_fence_after();
            {   // T_h -> cols [0,32), T_w -> cols [32,64): Q . table^T, N = 32 (27 table rows + zero fill)
                constexpr uint32_t idesc = umma_idesc_bf16(128, 32);
                const uint64_t dh = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RHA));
                const uint64_t dw = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_RWA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dqa + 2 * k, dh + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RHB)), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 32, dqa + 2 * k, dw + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base + 32, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_RWB)), idesc, 1u);
                umma_commit(t_full);
            }
            mbar_wait(k_full, 0);
            mbar_wait(t_read, 0);  // the softmax warps have taken T_h / T_w out of the columns S overwrites
            tc_fence_after();
            {   // S = Q K^T, N = 208
                constexpr uint32_t idesc = umma_idesc_bf16(128, WN_KEYS);
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(smem + WN_OFF_KA));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dqa + 2 * k, dk + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_bf16(tmem_base, dqb, umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_KB)), idesc, 1u);
                umma_commit(s_full);
            }
            mbar_wait(v_full, 0);
            mbar_wait(p_full, 0);
            tc_fence_after();
            {   // O = P V over 13 key steps of 16
                constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64), idesc16 = umma_idesc_bf16_bmn(128, 16);
#pragma unroll
                for (int ks = 0; ks < WN_KEYS / 16; ++ks) {
                    const uint64_t dp = ks < 12 ? umma_desc_sw128_kmajor(smem_u32(smem + (ks >> 2) * 16384)) + 2 * (ks & 3)
                                                : umma_desc_sw32_kmajor(smem_u32(smem + WN_OFF_P3));
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    umma_bf16(tmem_base, dp, umma_desc_sw128_mnmajor(smem_u32(smem + WN_OFF_VA + ks * 2048)), idesc64, acc);
                    umma_bf16(tmem_base + 64, dp, umma_desc_sw32_mnmajor(smem_u32(smem + WN_OFF_VB + ks * 512)), idesc16, acc);
                }
                umma_commit(o_full);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps: two threads per query row
        // warps 0-3 (half 0) own key columns [0,96), warps 4-7 (half 1) own [96,208); warp w and w+4 share a TMEM lane quadrant.
        const int quad = warp & 3, half = warp >> 2;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const int qtok = q0 + r;
        const bool q_ok = qtok < WN_S;
        const int qy = q_ok ? qtok / WN_KW : 0, qx = q_ok ? qtok - (qtok / WN_KW) * WN_KW : 0;
        mbar_wait(t_full, 0);
        tc_fence_after();
        {   // half 0 files T_h, half 1 files T_w
            uint32_t tt[32];
            tmem_ld_32x32(lane_addr + half * 32, tt);
            tmem_ld_wait();
            bf16* dst = half == 0 ? th_s : tw_s;
#pragma unroll
            for (int i = 0; i < 28; ++i) dst[i * AT_BQ + r] = __float2bfloat16_rn(__uint_as_float(tt[i]));
        }
        tc_fence_before();
        named_bar_sync(1, 256);
        if (lane == 0) mbar_arrive(t_read);
        float rh[WN_KW], rw[WN_KW];  // this row's rel_h[ky], rel_w[kx] (bf16 values), pre-multiplied by log2(e)
#pragma unroll
        for (int i = 0; i < WN_KW; ++i) {
            rh[i] = __bfloat162float(th_s[(qy - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
            rw[i] = __bfloat162float(tw_s[(qx - i + WN_KW - 1) * AT_BQ + r]) * AT_LOG2E;
        }
        mbar_wait(s_full, 0);
        tc_fence_after();
        // ---- pass 1: x = s*scale + bias (log2 domain) written back in place, row max over this thread's columns
        float mx = -INFINITY;
        uint32_t ra[32];
        uint32_t rc[16];
#define WINH_PASS1(C0)                                                  \
        {                                                               \
            tmem_ld_32x32(lane_addr + C0, ra);                          \
            tmem_ld_wait();                                             \
            mx = win_scores<C0, 32>(ra, rh, rw, p.scale_log2, mx);      \
            tmem_st_32x32(lane_addr + C0, ra);                          \
        }
        if (half == 0) {
            WINH_PASS1(0) WINH_PASS1(32) WINH_PASS1(64)
        } else {
            WINH_PASS1(96) WINH_PASS1(128) WINH_PASS1(160)
            tmem_ld_32x16(lane_addr + 192, rc);
            tmem_ld_wait();
            mx = win_scores<192, 16>(rc, rh, rw, p.scale_log2, mx);
            tmem_st_32x16(lane_addr + 192, rc);
        }
#undef WINH_PASS1
        tmem_st_wait();
        xbuf[half * AT_BQ + r] = mx;
        named_bar_sync(2 + quad, 64);
        mx = fmaxf(mx, xbuf[(half ^ 1) * AT_BQ + r]);
        // ---- pass 2: p = 2^(x - max) -> bf16 -> swizzled P (aliases the dead Q / K tiles)
        float sum = 0.f;
        auto exp_chunk = [&](const uint32_t(&raw)[32], int c0) {
            uint8_t* row = smem + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p0 = ex2_approx(__uint_as_float(raw[j8 * 8 + 2 * e]) - mx);
                    const float p1 = ex2_approx(__uint_as_float(raw[j8 * 8 + 2 * e + 1]) - mx);
                    sum += p0 + p1;
                    pk[e] = pack_bf16x2(p0, p1);
                }
                const int c16 = ((c0 & 63) >> 3) + j8;
                *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        };
#define WINH_PASS2(C0)                                  \
        {                                               \
            tmem_ld_32x32(lane_addr + C0, ra);          \
            tmem_ld_wait();                             \
            exp_chunk(ra, C0);                          \
        }
        if (half == 0) {
            WINH_PASS2(0) WINH_PASS2(32) WINH_PASS2(64)
        } else {
            WINH_PASS2(96) WINH_PASS2(128) WINH_PASS2(160)
            tmem_ld_32x16(lane_addr + 192, rc);
            tmem_ld_wait();
            uint8_t* row = smem + WN_OFF_P3 + r * 32;
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p0 = ex2_approx(__uint_as_float(rc[j8 * 8 + 2 * e]) - mx);
                    const float p1 = ex2_approx(__uint_as_float(rc[j8 * 8 + 2 * e + 1]) - mx);
                    sum += p0 + p1;
                    pk[e] = pack_bf16x2(p0, p1);
                }
                *reinterpret_cast<uint4*>(row + ((j8 ^ ((r >> 2) & 1)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
#undef WINH_PASS2
        xbuf[(2 + half) * AT_BQ + r] = sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- epilogue
        mbar_wait(o_full, 0);
        tc_fence_after();
        named_bar_sync(2 + quad, 64);
        const float tot = xbuf[2 * AT_BQ + r] + xbuf[3 * AT_BQ + r];
        const float inv = tot > 0.f ? 1.f / tot : 0.f;
        long long orow_i = q_ok ? (long long)(row_base + qtok) : -1;
        if (q_ok && p.out_map != nullptr) orow_i = p.out_map[row_base + qtok];
        const bool st_ok = orow_i >= 0;
        bf16* orow = p.out + (st_ok ? orow_i : 0) * p.out_ld + h * AT_HD;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int c0 = half * 48 + cc * 16;
            if (c0 < AT_HD) {
                uint32_t o[16];
                tmem_ld_32x16(lane_addr + c0, o);
                tmem_ld_wait();
                if (st_ok) {
                    uint4 u0, u1;
                    u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                    u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                    u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                    u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                    u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                    u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                    u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                    u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                    reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                    reinterpret_cast<uint4*>(orow + c0)[1] = u1;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}


// ------------------------------------------------------------------------------------------------ plain / causal attention
// softmax(q k^T * scale [+ causal mask]) v for head_dim 64 (CLIP ViT-L, 257 tokens) and 128 (LLaMA prefill, causal): the
// structure of sam_attn_tcgen05_kernel without the rel-pos prologue.  One CTA per (128-query tile, head, sequence); warp 0 = TMA
// producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = softmax (one query row per thread) / correction / epilogue.
//   smem  Q [HD/64 tiles of 128 x 64, 128B-swizzled] | 2 stages x (K, V: HD/64 tiles each) | P [2 tiles of 128 x 64 keys]
//   TMEM  S double buffer [0,128) [128,256), O [256, 256 + HD)
// q / k / v are row-major [tokens, heads*HD] matrices (any pitch) whose sequences follow each other, described by three 2-D tensor
// maps; rows of the next sequence (or past the end: zero-filled) that land in a tile are masked by key index / never stored.
// Replaces the mma.sync flash_attn_kernel on this path (70-95 TFLOP/s).
constexpr int FT_BQ = 128, FT_BK = 128, FT_NS = 2, FT_THREADS = 256, FT_O_COL = 256;
template <int HD> struct FtCfg {
    static constexpr int NC = HD / 64;                       // 64-column chunks of the head dimension
    static constexpr int Q_BYTES = NC * FT_BQ * 128;
    static constexpr int KV_BYTES = NC * FT_BK * 128;        // K or V of one stage
    static constexpr int STAGE = 2 * KV_BYTES;
    static constexpr int P_BYTES = FT_BQ * FT_BK * 2;
    static constexpr int SMEM = Q_BYTES + FT_NS * STAGE + P_BYTES + 1024 + 256;
};
struct FtParams {
    bf16* out;
    long long o_bs, o_ts, o_hs;
    int Sq, Sk;
    float scale_log2;
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(FT_THREADS, 1)
flash_attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const FtParams p) {
    using Cfg = FtCfg<HD>;
    constexpr int NC = Cfg::NC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem;
    uint8_t* stage0 = q_s + Cfg::Q_BYTES;
    auto k_s = [&](int s) { return stage0 + s * Cfg::STAGE; };
    auto v_s = [&](int s) { return stage0 + s * Cfg::STAGE + Cfg::KV_BYTES; };
    uint8_t* p_s = stage0 + FT_NS * Cfg::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + Cfg::P_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // FT_NS
    uint64_t* kv_empty = bars + 3;      // FT_NS
    uint64_t* s_full = bars + 5;        // 2
    uint64_t* s_empty = bars + 7;       // 2
    uint64_t* p_full = bars + 9;        // 1
    uint64_t* pv_done = bars + 10;      // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int off = p.Sk - p.Sq;                     // causal offset: query i sees keys <= i + off
    int kv_end = p.Sk;
    if (CAUSAL) kv_end = min(p.Sk, q0 + FT_BQ + off);
    const int n_tiles = (kv_end + FT_BK - 1) / FT_BK;
    const int q_row = b * p.Sq + q0, k_row = b * p.Sk;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < FT_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
            for (int c = 0; c < NC; ++c) tma_load_2d(q_s + c * (FT_BQ * 128), &tmQ, q_full, h * HD + c * 64, q_row);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % FT_NS;
                if (j >= FT_NS) mbar_wait(&kv_empty[s], ((j / FT_NS) - 1) & 1);
                mbar_arrive_expect_tx(&kv_full[s], Cfg::STAGE);
                const int r = k_row + j * FT_BK;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    tma_load_2d(k_s(s) + c * (FT_BK * 128), &tmK, &kv_full[s], h * HD + c * 64, r);
                    tma_load_2d(v_s(s) + c * (FT_BK * 128), &tmV, &kv_full[s], h * HD + c * 64, r);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            mbar_wait(q_full, 0);
            auto issue_s = [&](int j) {
                const int s = j % FT_NS, sb = j & 1;
                mbar_wait(&kv_full[s], (j / FT_NS) & 1);
                if (j >= 2) mbar_wait(&s_empty[sb], ((j >> 1) - 1) & 1);
                tc_fence_after();
                constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(q_s + c * (FT_BQ * 128)));
                    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(k_s(s) + c * (FT_BK * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + sb * FT_BK, da + 2 * k, db + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&s_full[sb]);
            };
            issue_s(0);
            constexpr uint32_t idesc64 = umma_idesc_bf16_bmn(128, 64);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                const int s = j % FT_NS;
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < FT_BK / 16; ++ks) {
                    const uint64_t dp = umma_desc_sw128_kmajor(smem_u32(p_s + (ks >> 2) * (FT_BQ * 128))) + 2 * (ks & 3);
                    const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        umma_bf16(tmem_base + FT_O_COL + c * 64, dp, umma_desc_sw128_mnmajor(smem_u32(v_s(s) + c * (FT_BK * 128) + ks * 2048)),
                                  idesc64, acc);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax / correction / epilogue
        const int quad = warp & 3;
        const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
        const int qtok = q0 + r;
        const int k_last = CAUSAL ? min(qtok + off, p.Sk - 1) : p.Sk - 1;   // last key this row attends to
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t xr[FT_BK];
#pragma unroll
            for (int c0 = 0; c0 < FT_BK; c0 += 32)
                tmem_ld_32x32(lane_addr + sb * FT_BK + c0, reinterpret_cast<uint32_t(&)[32]>(xr[c0]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);
            float x[FT_BK];
            float mx = -INFINITY;
            const int kbase = j * FT_BK;
            const uint64_t sc2 = f32x2_pack(p.scale_log2, p.scale_log2);   // two scores per FMUL2 / FADD2 (issue-bound loop)
#pragma unroll
            for (int c = 0; c < FT_BK; c += 2) {
                float x0, x1;
                f32x2_unpack(f32x2_mul(f32x2_pack(__uint_as_float(xr[c]), __uint_as_float(xr[c + 1])), sc2), x0, x1);
                x[c] = (kbase + c <= k_last) ? x0 : -INFINITY;
                x[c + 1] = (kbase + c + 1 <= k_last) ? x1 : -INFINITY;
                mx = fmaxf(mx, fmaxf(x[c], x[c + 1]));
            }
            // online softmax with lazy rescaling (see sam_attn_tcgen05_kernel); a row whose keys are all masked keeps m_ref
            float corr = 1.f;
            bool moved = false;
            if (m_ref == -INFINITY) {
                m_ref = mx;          // first tile with a live key (tile 0 always has one for rows < Sq)
            } else if (mx > m_ref + 8.f) {
                corr = ex2_approx(m_ref - mx);
                m_ref = mx;
                moved = true;
            }
            const float mr = (m_ref == -INFINITY) ? 0.f : m_ref;
            const uint64_t nmr2 = f32x2_pack(-mr, -mr);
            uint64_t sum2 = f32x2_pack(0.f, 0.f);
            uint32_t pk[FT_BK / 2];
#pragma unroll
            for (int c = 0; c < FT_BK; c += 2) {
                float d0, d1;
                f32x2_unpack(f32x2_add(f32x2_pack(x[c], x[c + 1]), nmr2), d0, d1);
                const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
                sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                pk[c >> 1] = pack_bf16x2(p0, p1);
            }
            float sum_lo, sum_hi;
            f32x2_unpack(sum2, sum_lo, sum_hi);
            l_run = l_run * corr + (sum_lo + sum_hi);
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, moved)) {
#pragma unroll
                    for (int c0 = 0; c0 < HD; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld_32x16(lane_addr + FT_O_COL + c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st_32x16(lane_addr + FT_O_COL + c0, o);
                    }
                    tmem_st_wait();
                }
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                uint8_t* row = p_s + ch * (FT_BQ * 128) + r * 128;
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) {
                    const int i = ch * 32 + c16 * 4;
                    *reinterpret_cast<uint4*>(row + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[i], pk[i + 1], pk[i + 2], pk[i + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* orow = p.out + (long long)b * p.o_bs + (long long)qtok * p.o_ts + (long long)h * p.o_hs;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_32x16(lane_addr + FT_O_COL + c0, o);
            tmem_ld_wait();
            if (qtok < p.Sq) {
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(orow + c0)[0] = u0;
                reinterpret_cast<uint4*>(orow + c0)[1] = u1;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD, bool CAUSAL>
static int launch_ft(ivlm_ctx* h, const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv, const FtParams& p, dim3 grid,
                     cudaStream_t stream) {
    using Cfg = FtCfg<HD>;
    const uint64_t bit = 1ull << (24 + (HD == 128 ? 0 : 2) + (CAUSAL ? 1 : 0));
    if (!(h->attr_done & bit)) {
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_tcgen05_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        h->attr_done |= bit;
    }
    flash_attn_tcgen05_kernel<HD, CAUSAL><<<grid, FT_THREADS, Cfg::SMEM, stream>>>(*tq, *tk, *tv, p);
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

// Returns 1 when the launch was taken by the tcgen05 kernel, 0 when the layout does not fit it (the caller falls back to
// flash_attn_kernel), < 0 on error.
int attention_tcgen05_try(ivlm_ctx* h, const ivlm_attn_args* a, cudaStream_t stream) {
    if (h->attn_variant == 1 || a->rel_h != nullptr || (a->D != 64 && a->D != 128)) return 0;
    // sequences follow each other with one row pitch, heads are HD apart, 16-byte aligned bases / pitches / outputs
    if (a->q_hs != a->D || a->k_hs != a->D || a->v_hs != a->D) return 0;
    if (a->q_bs != (int64_t)a->Sq * a->q_ts || a->k_bs != (int64_t)a->Sk * a->k_ts || a->v_bs != (int64_t)a->Sk * a->v_ts) return 0;
    if (a->q_ts % 8 || a->k_ts % 8 || a->v_ts % 8 || a->o_ts % 8 || a->o_hs % 8 || a->o_bs % 8) return 0;
    if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
         reinterpret_cast<uintptr_t>(a->out)) & 15) return 0;
    const uint64_t cols = (uint64_t)a->H * a->D;
    const CUtensorMap *tq, *tk, *tv;
    if (get_tmap_bf16_ex(h, a->q, (uint64_t)a->B * a->Sq, cols, (uint64_t)a->q_ts, FT_BQ, 64, 128, &tq) != IVLM_OK) return -1;
    if (get_tmap_bf16_ex(h, a->k, (uint64_t)a->B * a->Sk, cols, (uint64_t)a->k_ts, FT_BK, 64, 128, &tk) != IVLM_OK) return -1;
    if (get_tmap_bf16_ex(h, a->v, (uint64_t)a->B * a->Sk, cols, (uint64_t)a->v_ts, FT_BK, 64, 128, &tv) != IVLM_OK) return -1;
    FtParams p;
    p.out = reinterpret_cast<bf16*>(a->out);
    p.o_bs = a->o_bs; p.o_ts = a->o_ts; p.o_hs = a->o_hs;
    p.Sq = a->Sq; p.Sk = a->Sk;
    p.scale_log2 = a->scale * AT_LOG2E;
    dim3 grid((a->Sq + FT_BQ - 1) / FT_BQ, a->H, a->B);
    int rc;
    if (a->D == 128) rc = a->causal ? launch_ft<128, true>(h, tq, tk, tv, p, grid, stream) : launch_ft<128, false>(h, tq, tk, tv, p, grid, stream);
    else rc = a->causal ? launch_ft<64, true>(h, tq, tk, tv, p, grid, stream) : launch_ft<64, false>(h, tq, tk, tv, p, grid, stream);
    return rc == IVLM_OK ? 1 : -1;
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_sam_attention_bf16(ivlm_handle h, const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                                       void* out, int32_t B, int32_t heads, int32_t Hq, int32_t Wq, int32_t hd,
                                       int64_t out_ld, const int32_t* out_row_map, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && qkv && rel_pos_h && rel_pos_w && out && B > 0 && heads > 0, "sam_attention: bad arguments");
    IVLM_REQUIRE(hd == AT_HD, "sam_attention: head_dim %d not instantiated (80)", hd);
    IVLM_REQUIRE((Hq == 64 && Wq == 64) || (Hq == 14 && Wq == 14), "sam_attention: token grid %dx%d not instantiated (64x64, 14x14)",
                 Hq, Wq);
    IVLM_REQUIRE(out_ld % 8 == 0, "sam_attention: out pitch must be a multiple of 8 elements");
    const int S = Hq * Wq, E = heads * hd;
    const uint64_t rows = (uint64_t)B * S;
    const CUtensorMap *qa, *qb, *rha, *rhb, *rwa, *rwb;
    IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, AT_BQ, 64, 128, &qa));
    IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, AT_BQ, 16, 32, &qb));
    IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_h, 2 * Hq - 1, hd, hd, AT_BK, 64, 128, &rha));
    IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_h, 2 * Hq - 1, hd, hd, AT_BK, 16, 32, &rhb));
    IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_w, 2 * Wq - 1, hd, hd, AT_BK, 64, 128, &rwa));
    IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_w, 2 * Wq - 1, hd, hd, AT_BK, 16, 32, &rwb));
    SamAttnParams p;
    p.out = reinterpret_cast<bf16*>(out);
    p.out_ld = out_ld;
    p.S = S;
    p.heads = heads;
    p.KH = Hq;
    p.scale_log2 = (1.0f / sqrtf((float)hd)) * AT_LOG2E;
    p.out_map = out_row_map;
    p.prefetch_ahead = h->attn_prefetch_ahead;
    IVLM_REQUIRE(out_row_map == nullptr || (Wq == 14 && h->window_attn_variant != 1),
                 "sam_attention: out_row_map is implemented by the 14x14 window kernel only");
    dim3 grid((S + AT_BQ - 1) / AT_BQ, heads, B);
    if (!(h->attr_done & (1ull << 16))) {   // per handle = per device
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_tcgen05_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_tcgen05_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_window_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WN_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_window_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WNH_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_window_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WN_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_global64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_attn_global64h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2H_SMEM));
        h->attr_done |= 1ull << 16;
    }
    if (Wq == 64 && (h->global_attn_variant == 0 || h->global_attn_variant == 2)) {
        const CUtensorMap *kva, *kvb;
        IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, G2_BK, 64, 128, &kva));
        IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, G2_BK, 16, 32, &kvb));
        if (h->global_attn_variant == 0)
            sam_attn_global64h_kernel<<<grid, G2H_THREADS, G2H_SMEM, stream>>>(*qa, *qb, *kva, *kvb, *rha, *rhb, *rwa, *rwb, p);
        else
            sam_attn_global64_kernel<<<grid, G2_THREADS, G2_SMEM, stream>>>(*qa, *qb, *kva, *kvb, *rha, *rhb, *rwa, *rwb, p);
    } else if (Wq == 64) {
        sam_attn_tcgen05_kernel<64><<<grid, AT_THREADS, AT_SMEM, stream>>>(*qa, *qb, *rha, *rhb, *rwa, *rwb, p);
    } else if (h->window_attn_variant == 0 || h->window_attn_variant == 2 || h->window_attn_variant == 3) {
        const CUtensorMap *kva, *kvb, *wha, *whb, *wwa, *wwb;
        IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, WN_KEYS, 64, 128, &kva));
        IVLM_TRY(get_tmap_bf16_ex(h, qkv, rows, 3 * (uint64_t)E, 3 * (uint64_t)E, WN_KEYS, 16, 32, &kvb));
        IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_h, 2 * Hq - 1, hd, hd, 32, 64, 128, &wha));
        IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_h, 2 * Hq - 1, hd, hd, 32, 16, 32, &whb));
        IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_w, 2 * Wq - 1, hd, hd, 32, 64, 128, &wwa));
        IVLM_TRY(get_tmap_bf16_ex(h, rel_pos_w, 2 * Wq - 1, hd, hd, 32, 16, 32, &wwb));
        if (h->window_attn_variant == 3) {
            const int n_items = 2 * heads * B;
            const int ctas = n_items < 2 * h->num_sms ? n_items : 2 * h->num_sms;
            sam_attn_window_persist_kernel<<<ctas, WN_THREADS, WN_SMEM, stream>>>(*qa, *qb, *kva, *kvb, *wha, *whb, *wwa, *wwb, p, n_items);
        } else if (h->window_attn_variant == 2)   // measured 201 vs 213 TFLOP/s: the window kernel waits on loads, not on its softmax warps
            sam_attn_window_h_kernel<<<grid, WNH_THREADS, WNH_SMEM, stream>>>(*qa, *qb, *kva, *kvb, *wha, *whb, *wwa, *wwb, p);
        else
            sam_attn_window_tcgen05_kernel<<<grid, WN_THREADS, WN_SMEM, stream>>>(*qa, *qb, *kva, *kvb, *wha, *whb, *wwa, *wwb, p);
    } else {
        sam_attn_tcgen05_kernel<14><<<grid, AT_THREADS, AT_SMEM, stream>>>(*qa, *qb, *rha, *rhb, *rwa, *rwb, p);
    }
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}
