// Attention kernels.
//  * flash_attn_kernel: fused softmax(QK^T*scale + bias)V, online softmax in fp32, bf16 mma.sync m16n8k16
//    fragments fed by ldmatrix from cp.async double-buffered K/V tiles. Covers SAM window (14x14) and global
//    (64x64) attention with the decomposed relative-position bias, CLIP (257 tokens) and LLaMA prefill (causal).
//    (Round-1 implementation on the legacy tensor path; attention is ~8% of the path's FLOPs, the dense
//    projections run on tcgen05 in gemm_tcgen05.cu.)
//  * sam_relpos_kernel: the two decomposed rel-pos terms q.Rh / q.Rw.
//  * decode_attn_paged_kernel: one-token attention over the paged KV cache (HBM-bound, warp-shuffle dots).
//  * attn_few_queries / attn_few_keys: the SAM two-way decoder's 9-token attentions (warp-shuffle reductions).
#include "common.cuh"
#include "runtime.h"

namespace ivlm {

struct AttnParams {
    const bf16 *q, *k, *v;
    bf16* out;
    long long q_bs, q_ts, q_hs, k_bs, k_ts, k_hs, v_bs, v_ts, v_hs, o_bs, o_ts, o_hs;
    int B, H, Sq, Sk;
    float scale;
    const float *rel_h, *rel_w;
    int kh, kw;
};

constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

template <int HD>
IVLM_DEVINL void fa_load_tile(bf16* dst, const bf16* base, long long ts, int row0, int limit) {
    constexpr int LDS = HD + 8;
    constexpr int CPR = HD / 8;
    for (int i = threadIdx.x; i < 64 * CPR; i += FA_THREADS) {
        const int r = i / CPR, c = i % CPR;
        const int gr = row0 + r;
        const bool ok = gr < limit;
        const bf16* src = base + (long long)(ok ? gr : 0) * ts + c * 8;
        cp_async_16(dst + r * LDS + c * 8, src, ok);
    }
}

template <int HD, bool CAUSAL, bool RELPOS>
__global__ void __launch_bounds__(FA_THREADS) flash_attn_kernel(const AttnParams p) {
    constexpr int LDS = HD + 8;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    bf16* Qs = reinterpret_cast<bf16*>(fa_smem);
    bf16* Ks = Qs + 64 * LDS;
    bf16* Vs = Ks + 2 * 64 * LDS;
    float* relh_s = reinterpret_cast<float*>(Vs + 2 * 64 * LDS);
    float* relw_s = relh_s + (RELPOS ? 64 * p.kh : 0);

    const int q0 = blockIdx.x * FA_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    const bf16* qb = p.q + b * p.q_bs + h * p.q_hs;
    const bf16* kb = p.k + b * p.k_bs + h * p.k_hs;
    const bf16* vb = p.v + b * p.v_bs + h * p.v_hs;
    const int off = p.Sk - p.Sq;  // causal offset

    int kv_end = p.Sk;
    if (CAUSAL) kv_end = min(p.Sk, q0 + FA_BQ + off);
    const int n_tiles = (kv_end + FA_BK - 1) / FA_BK;

    fa_load_tile<HD>(Qs, qb, p.q_ts, q0, p.Sq);
    fa_load_tile<HD>(Ks, kb, p.k_ts, 0, p.Sk);
    fa_load_tile<HD>(Vs, vb, p.v_ts, 0, p.Sk);
    cp_async_commit();
    if (RELPOS) {
        const long long base = ((long long)b * p.H + h) * p.Sq;
        for (int i = threadIdx.x; i < 64 * p.kh; i += FA_THREADS) {
            const int r = i / p.kh, c = i % p.kh;
            relh_s[i] = (q0 + r < p.Sq) ? p.rel_h[(base + q0 + r) * p.kh + c] : 0.f;
        }
        for (int i = threadIdx.x; i < 64 * p.kw; i += FA_THREADS) {
            const int r = i / p.kw, c = i % p.kw;
            relw_s[i] = (q0 + r < p.Sq) ? p.rel_w[(base + q0 + r) * p.kw + c] : 0.f;
        }
    }

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int rl0 = warp * 16 + g, rl1 = rl0 + 8;  // local query rows of this thread
    const int row0 = q0 + rl0, row1 = q0 + rl1;

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) {
            fa_load_tile<HD>(Ks + (buf ^ 1) * 64 * LDS, kb, p.k_ts, (j + 1) * FA_BK, p.Sk);
            fa_load_tile<HD>(Vs + (buf ^ 1) * 64 * LDS, vb, p.v_ts, (j + 1) * FA_BK, p.Sk);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const bf16* Kt = Ks + buf * 64 * LDS;
        const bf16* Vt = Vs + buf * 64 * LDS;

        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t a[4];
            ldmatrix_x4(a[0], a[1], a[2], a[3], Qs + (warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3,
                            Kt + (n2 * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16_16816(s[2 * n2], a, b0, b1);
                mma_bf16_16816(s[2 * n2 + 1], a, b2, b3);
            }
        }
        // ---- scale, bias, mask, online softmax
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = j * FA_BK + nt * 8 + 2 * t + (e & 1);
                const int row = (e < 2) ? row0 : row1;
                float x = s[nt][e] * p.scale;
                if (RELPOS) {
                    const int rl = (e < 2) ? rl0 : rl1;
                    const int ch = col / p.kw, cw = col - ch * p.kw;
                    if (col < p.Sk) x += relh_s[rl * p.kh + ch] + relw_s[rl * p.kw + cw];
                }
                bool dead = col >= p.Sk;
                if (CAUSAL) dead = dead || (col > row + off);
                x = dead ? -INFINITY : x * LOG2E;
                s[nt][e] = x;
                mx[e >> 1] = fmaxf(mx[e >> 1], x);
            }
        }
        float corr[2], msafe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            msafe[r] = (m_new == -INFINITY) ? 0.f : m_new;
            corr[r] = exp2f(m_run[r] - msafe[r]);  // m_run = -inf -> 0
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pa[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - msafe[0]), p1 = exp2f(s[nt][1] - msafe[0]);
            const float p2 = exp2f(s[nt][2] - msafe[1]), p3 = exp2f(s[nt][3] - msafe[1]);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            l_run[r] = l_run[r] * corr[r] + rs[r];
        }
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int n2 = 0; n2 < HD / 16; ++n2) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(b0, b1, b2, b3,
                                  Vt + (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + n2 * 16 + (lane >> 4) * 8);
                mma_bf16_16816(o[2 * n2], pa[kk], b0, b1);
                mma_bf16_16816(o[2 * n2 + 1], pa[kk], b2, b3);
            }
        }
        __syncthreads();
    }

    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    bf16* ob = p.out + b * p.o_bs + h * p.o_hs;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        const int c = i * 8 + 2 * t;
        if (row0 < p.Sq)
            *reinterpret_cast<uint32_t*>(ob + (long long)row0 * p.o_ts + c) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < p.Sq)
            *reinterpret_cast<uint32_t*>(ob + (long long)row1 * p.o_ts + c) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
    }
}

template <int HD, bool CAUSAL, bool RELPOS>
static int launch_fa(ivlm_ctx* h, const AttnParams& p, cudaStream_t stream) {
    constexpr int LDS = HD + 8;
    size_t smem = (size_t)(64 + 4 * 64) * LDS * 2;
    if (RELPOS) smem += (size_t)64 * (p.kh + p.kw) * 4;
    IVLM_REQUIRE(smem <= 200 * 1024, "attention: rel-pos tables too large for shared memory (kh=%d kw=%d)", p.kh, p.kw);
    IVLM_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, CAUSAL, RELPOS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.Sq + FA_BQ - 1) / FA_BQ, p.H, p.B);
    flash_attn_kernel<HD, CAUSAL, RELPOS><<<grid, FA_THREADS, smem, stream>>>(p);
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

// ------------------------------------------------------------------------------------------------ SAM rel-pos
// CTA per (b, head, query row qy). rel_h[b,h,(qy,qx),kh] = q . Rh[qy-kh+Hk-1]; rel_w[...,kw] = q . Rw[qx-kw+Wk-1].
// The reference evaluates both einsums in bf16 (image_encoder.py:383-384), so outputs carry bf16-rounded values.
__global__ void sam_relpos_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ rph,
                                  const bf16* __restrict__ rpw, float* __restrict__ rel_h, float* __restrict__ rel_w,
                                  int heads, int Hq, int Wq, int hd) {
    extern __shared__ __align__(16) uint8_t rp_smem[];
    const int LD = hd + 2;
    bf16* q_s = reinterpret_cast<bf16*>(rp_smem);  // [Wq][LD]
    bf16* rh_s = q_s + Wq * LD;                    // [Hq][LD]   rows kh -> Rh[qy-kh+Hq-1]
    bf16* rw_s = rh_s + Hq * LD;                   // [2Wq-1][LD]
    const int qy = blockIdx.x, hh = blockIdx.y, b = blockIdx.z;
    const int S = Hq * Wq;
    const long long ld_qkv = 3LL * heads * hd;
    for (int i = threadIdx.x; i < Wq * hd; i += blockDim.x) {
        const int qx = i / hd, c = i % hd;
        q_s[qx * LD + c] = qkv[((long long)b * S + qy * Wq + qx) * ld_qkv + hh * hd + c];
    }
    for (int i = threadIdx.x; i < Hq * hd; i += blockDim.x) {
        const int kh = i / hd, c = i % hd;
        rh_s[kh * LD + c] = rph[(qy - kh + Hq - 1) * hd + c];
    }
    for (int i = threadIdx.x; i < (2 * Wq - 1) * hd; i += blockDim.x) {
        const int r = i / hd, c = i % hd;
        rw_s[r * LD + c] = rpw[r * hd + c];
    }
    __syncthreads();
    const long long obase = (((long long)b * heads + hh) * S + qy * Wq);
    for (int i = threadIdx.x; i < Wq * Hq; i += blockDim.x) {
        const int qx = i / Hq, kh = i % Hq;
        float acc = 0.f;
        for (int c = 0; c < hd; ++c) acc += __bfloat162float(q_s[qx * LD + c]) * __bfloat162float(rh_s[kh * LD + c]);
        rel_h[(obase + qx) * Hq + kh] = bf16_round(acc);
    }
    for (int i = threadIdx.x; i < Wq * Wq; i += blockDim.x) {
        const int qx = i / Wq, kw = i % Wq;
        const bf16* r = rw_s + (qx - kw + Wq - 1) * LD;
        float acc = 0.f;
        for (int c = 0; c < hd; ++c) acc += __bfloat162float(q_s[qx * LD + c]) * __bfloat162float(r[c]);
        rel_w[(obase + qx) * Wq + kw] = bf16_round(acc);
    }
}

// ------------------------------------------------------------------------------------------------ paged decode
// CTA per (b, head), 8 warps. HF LlamaAttention eager numerics (transformers 4.31): scores = bf16(bf16(q.k)/sqrt(hd)),
// softmax in fp32 cast to bf16, bf16 P.V.  HBM-bound.  The cache is [pages, H, 16, HD]: one (page, head) is a contiguous
// 16*HD*2-byte run, and one warp iteration consumes exactly one page with 16/RPI independent 16-byte loads per lane
// (a single block-table lookup, no per-row index arithmetic).
constexpr int DEC_PAGE = 16, DEC_MAX_PAGES = 256;
template <int HD, int DEC_WARPS, bool DEC_L2_PREFETCH>
__global__ void __launch_bounds__(DEC_WARPS * 32) decode_attn_paged_kernel(const bf16* __restrict__ q,
                                                                          const bf16* __restrict__ k_cache,
                                                                          const bf16* __restrict__ v_cache,
                                                                          const int* __restrict__ block_table,
                                                                          const int* __restrict__ seq_lens,
                                                                          bf16* __restrict__ out, int H, int max_pages,
                                                                          float inv_scale) {
    extern __shared__ float sc[];  // [seq_len] scores, then probabilities
    __shared__ float red[DEC_WARPS];
    constexpr int LPR = HD / 8;            // lanes per row (16 for HD=128, 8 for HD=64)
    constexpr int RPI = 32 / LPR;          // rows per warp-wide load instruction
    constexpr int NLD = DEC_PAGE / RPI;    // load instructions per page
    __shared__ float part[DEC_WARPS * RPI][HD];
    __shared__ int bt[DEC_MAX_PAGES];
    pdl_wait_then_launch();
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, li = lane % LPR;
    const int len = seq_lens[b];
    const int n_pages = (len + DEC_PAGE - 1) / DEC_PAGE;
    // the whole row of the block table (it exists up to max_pages): no dependency on the seq_lens load above
    for (int i = threadIdx.x; i < max_pages; i += DEC_WARPS * 32) bt[i] = block_table[(long long)b * max_pages + i];
    float qr[8];
    {
        const uint4 u = *reinterpret_cast<const uint4*>(q + ((long long)b * H + h) * HD + li * 8);
        const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y), d = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
        qr[0] = a.x; qr[1] = a.y; qr[2] = c.x; qr[3] = c.y; qr[4] = d.x; qr[5] = d.y; qr[6] = e.x; qr[7] = e.y;
    }
    __syncthreads();
    const long long lane_off = (long long)sub * HD + li * 8;  // this lane's element offset inside a (page, head) run
    // Experiment kept as an option (off): every K and V line this warp will touch is requested into L2 up front (one request per
    // 128-byte line) so that the page loops below -- one page per iteration, few registers, three CTAs resident per SM -- would not
    // pay an HBM round trip per iteration.  Measured 19.6 us against 16.0 us per layer: slower.
    if (DEC_L2_PREFETCH && (li & 7) == 0) {
        for (int pg = warp; pg < n_pages; pg += DEC_WARPS) {
            const long long o = ((long long)bt[pg] * H + h) * (DEC_PAGE * HD) + lane_off;
            const int k0 = pg * DEC_PAGE;
#pragma unroll
            for (int u = 0; u < NLD; ++u)
                if (k0 + u * RPI + sub < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(k_cache + o + u * RPI * HD));
#pragma unroll
            for (int u = 0; u < NLD; ++u)
                if (k0 + u * RPI + sub < len) asm volatile("prefetch.global.L2 [%0];" ::"l"(v_cache + o + u * RPI * HD));
        }
    }
    // ---- phase 1: scores.  The loops of both phases are software pipelined over HALF pages with two register buffers: the loads
    // of unit i+1 are in flight while unit i is reduced, so a warp keeps requests outstanding all the time.  With one page per
    // iteration all 320 CTAs issued a burst, waited out the HBM latency, computed, and issued the next burst (DRAM 40 % busy).
    constexpr int UL = NLD / 2;                 // loads per unit (half a page)
    const int my_pages = n_pages > warp ? (n_pages - warp + DEC_WARPS - 1) / DEC_WARPS : 0;
    const int nU = 2 * my_pages;
    uint4 bufA[UL], bufB[UL];
    auto unit_k0 = [&](int ui) { return (warp + (ui >> 1) * DEC_WARPS) * DEC_PAGE + (ui & 1) * UL * RPI; };
    auto unit_off = [&](int ui) {
        return ((long long)bt[warp + (ui >> 1) * DEC_WARPS] * H + h) * (DEC_PAGE * HD) + lane_off + (long long)(ui & 1) * UL * RPI * HD;
    };
    auto load_unit = [&](uint4(&buf)[UL], const bf16* cache, int ui) {
        const bf16* base = cache + unit_off(ui);
        const int k0 = unit_k0(ui);
#pragma unroll
        for (int u = 0; u < UL; ++u)
            buf[u] = (k0 + u * RPI + sub < len) ? *reinterpret_cast<const uint4*>(base + u * RPI * HD) : make_uint4(0, 0, 0, 0);
    };
    float lmax = -INFINITY;
    auto score_unit = [&](const uint4(&buf)[UL], int ui) {
        const int k0 = unit_k0(ui);
        float d[UL];
#pragma unroll
        for (int u = 0; u < UL; ++u) {
            const float2 a = unpack_bf16x2(buf[u].x), c = unpack_bf16x2(buf[u].y), e = unpack_bf16x2(buf[u].z), f = unpack_bf16x2(buf[u].w);
            d[u] = qr[0] * a.x + qr[1] * a.y + qr[2] * c.x + qr[3] * c.y + qr[4] * e.x + qr[5] * e.y + qr[6] * f.x + qr[7] * f.y;
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < UL; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
        }
#pragma unroll
        for (int u = 0; u < UL; ++u) {
            const int kpos = k0 + u * RPI + sub;
            if (kpos < len) {
                const float x = bf16_round(bf16_round(d[u]) / inv_scale);
                if (li == 0) sc[kpos] = x;
                lmax = fmaxf(lmax, x);
            }
        }
    };
    if (nU > 0) load_unit(bufA, k_cache, 0);
    for (int ui = 0; ui < nU; ui += 2) {     // nU is even
        load_unit(bufB, k_cache, ui + 1);
        score_unit(bufA, ui);
        if (ui + 2 < nU) load_unit(bufA, k_cache, ui + 2);
        score_unit(bufB, ui + 1);
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float gmax = red[0];
#pragma unroll
    for (int w = 1; w < DEC_WARPS; ++w) gmax = fmaxf(gmax, red[w]);
    __syncthreads();
    // ---- phase 2: softmax (fp32), probabilities rounded to bf16 like the reference's .to(query.dtype)
    float lsum = 0.f;
    for (int i = threadIdx.x; i < len; i += DEC_WARPS * 32) {
        const float e = __expf(sc[i] - gmax);
        sc[i] = e;
        lsum += e;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) tot += red[w];
    const float inv = 1.f / tot;
    // ---- phase 3: P.V; every (warp, sub-row) slot accumulates its keys, partials reduced through shared memory
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    uint64_t acc2[4] = {f32x2_pack(0.f, 0.f), f32x2_pack(0.f, 0.f), f32x2_pack(0.f, 0.f), f32x2_pack(0.f, 0.f)};
    auto pv_unit = [&](const uint4(&buf)[UL], int ui) {   // rows in cache order: the accumulation order of the unpipelined loop
        const int k0 = unit_k0(ui);
#pragma unroll
        for (int u = 0; u < UL; ++u) {
            const int kpos = k0 + u * RPI + sub;
            const float p_ = kpos < len ? bf16_round(sc[kpos] * inv) : 0.f;
            const float2 a = unpack_bf16x2(buf[u].x), c = unpack_bf16x2(buf[u].y), e = unpack_bf16x2(buf[u].z), f = unpack_bf16x2(buf[u].w);
            const uint64_t p2 = f32x2_pack(p_, p_);   // FFMA2: two of the eight independent accumulators per instruction, same fma per element
            acc2[0] = f32x2_fma(p2, f32x2_pack(a.x, a.y), acc2[0]);
            acc2[1] = f32x2_fma(p2, f32x2_pack(c.x, c.y), acc2[1]);
            acc2[2] = f32x2_fma(p2, f32x2_pack(e.x, e.y), acc2[2]);
            acc2[3] = f32x2_fma(p2, f32x2_pack(f.x, f.y), acc2[3]);
        }
    };
    if (nU > 0) load_unit(bufA, v_cache, 0);
    for (int ui = 0; ui < nU; ui += 2) {
        load_unit(bufB, v_cache, ui + 1);
        pv_unit(bufA, ui);
        if (ui + 2 < nU) load_unit(bufA, v_cache, ui + 2);
        pv_unit(bufB, ui + 1);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) f32x2_unpack(acc2[e], acc[2 * e], acc[2 * e + 1]);
#pragma unroll
    for (int e = 0; e < 8; ++e) part[warp * RPI + sub][li * 8 + e] = acc[e];
    __syncthreads();
    for (int d = threadIdx.x; d < HD; d += DEC_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < DEC_WARPS * RPI; ++w) s += part[w][d];
        out[((long long)b * H + h) * HD + d] = __float2bfloat16_rn(s);
    }
}

// ------------------------------------------------------------------------------------------------ SAM decoder attn
// Few queries (<=16), many keys: CTA per (b, head); per query one pass over the keys with block reductions.
// Rounding follows the eager bf16 reference (transformer.py:232-238): scores, scaled scores, probabilities and the
// output are each bf16-rounded.
template <int HD>
__global__ void __launch_bounds__(256) attn_few_queries_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                               const bf16* __restrict__ v, bf16* __restrict__ out,
                                                               int q_bcast, int Nq, int Nk, int heads) {
    const int hh = blockIdx.x, b = blockIdx.y;
    const int C = heads * HD;
    __shared__ float qs[HD];
    __shared__ float red[8][HD + 1];
    __shared__ float bc[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bf16* kb = k + (long long)b * Nk * C + hh * HD;
    const bf16* vb = v + (long long)b * Nk * C + hh * HD;
    const float inv_sqrt = sqrtf((float)HD);
    for (int qi = 0; qi < Nq; ++qi) {
        __syncthreads();
        if (threadIdx.x < HD)
            qs[threadIdx.x] = __bfloat162float(q[((long long)(q_bcast ? 0 : b) * Nq + qi) * C + hh * HD + threadIdx.x]);
        __syncthreads();
        // pass 1: max
        float lmax = -INFINITY;
        for (int j = threadIdx.x; j < Nk; j += 256) {
            const uint4* kp = reinterpret_cast<const uint4*>(kb + (long long)j * C);
            float d = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                uint4 u = kp[c8];
                float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                const float* qq = qs + c8 * 8;
                d += qq[0] * a.x + qq[1] * a.y + qq[2] * bb.x + qq[3] * bb.y + qq[4] * c.x + qq[5] * c.y + qq[6] * e.x +
                     qq[7] * e.y;
            }
            d = bf16_round(bf16_round(d) / inv_sqrt);
            lmax = fmaxf(lmax, d);
        }
        lmax = warp_max(lmax);
        if (lane == 0) red[warp][0] = lmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = red[0][0];
            for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][0]);
            bc[0] = m;
        }
        __syncthreads();
        const float gmax = bc[0];
        // pass 2: sum of exp
        float lsum = 0.f;
        for (int j = threadIdx.x; j < Nk; j += 256) {
            const uint4* kp = reinterpret_cast<const uint4*>(kb + (long long)j * C);
            float d = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                uint4 u = kp[c8];
                float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                const float* qq = qs + c8 * 8;
                d += qq[0] * a.x + qq[1] * a.y + qq[2] * bb.x + qq[3] * bb.y + qq[4] * c.x + qq[5] * c.y + qq[6] * e.x +
                     qq[7] * e.y;
            }
            d = bf16_round(bf16_round(d) / inv_sqrt);
            lsum += __expf(d - gmax);
        }
        lsum = warp_sum(lsum);
        __syncthreads();
        if (lane == 0) red[warp][0] = lsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w][0];
            bc[1] = 1.f / s;
        }
        __syncthreads();
        const float inv = bc[1];
        // pass 3: weighted sum of V
        float acc[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = 0.f;
        for (int j = threadIdx.x; j < Nk; j += 256) {
            const uint4* kp = reinterpret_cast<const uint4*>(kb + (long long)j * C);
            float d = 0.f;
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                uint4 u = kp[c8];
                float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                const float* qq = qs + c8 * 8;
                d += qq[0] * a.x + qq[1] * a.y + qq[2] * bb.x + qq[3] * bb.y + qq[4] * c.x + qq[5] * c.y + qq[6] * e.x +
                     qq[7] * e.y;
            }
            d = bf16_round(bf16_round(d) / inv_sqrt);
            const float pr = bf16_round(__expf(d - gmax) * inv);
            const uint4* vp = reinterpret_cast<const uint4*>(vb + (long long)j * C);
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                uint4 u = vp[c8];
                float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                float* ac = acc + c8 * 8;
                ac[0] += pr * a.x; ac[1] += pr * a.y; ac[2] += pr * bb.x; ac[3] += pr * bb.y;
                ac[4] += pr * c.x; ac[5] += pr * c.y; ac[6] += pr * e.x; ac[7] += pr * e.y;
            }
        }
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = warp_sum(acc[c]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < HD; ++c) red[warp][c] = acc[c];
        }
        __syncthreads();
        if (threadIdx.x < HD) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
            out[((long long)b * Nq + qi) * C + hh * HD + threadIdx.x] = __float2bfloat16_rn(s);
        }
    }
}

// All NQ queries of a (sample, head) in one sweep: the K slice is read ONCE (scores of every query per key row, parked in shared
// memory as the bf16 values the reference rounds them to), the softmax statistics come from shared memory, the V slice is read
// ONCE with NQ x 16 accumulators per thread.  The per-query form above reads K three times and V once PER QUERY (27 + 9 sweeps
// for the 9 decoder tokens: 0.33 ms per launch at 32 views); same rounding points, same per-thread key striding and reduction
// order (so the outputs are bit-identical to it).
template <int NQ>
__global__ void __launch_bounds__(256, 1) attn_few_queries_all_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                                      const bf16* __restrict__ v, bf16* __restrict__ out,
                                                                      int q_bcast, int Nq, int Nk, int heads) {
    constexpr int HD = 16;
    extern __shared__ __align__(16) uint8_t fq_smem[];
    bf16* sc = reinterpret_cast<bf16*>(fq_smem);                        // [NQ][Nk] scores, then probabilities
    float* red = reinterpret_cast<float*>(fq_smem + (size_t)NQ * Nk * 2);   // [8][NQ*HD]
    float* stat = red + 8 * NQ * HD;                                     // [NQ] max, [NQ] 1/sum
    const int hh = blockIdx.x, b = blockIdx.y;
    const int C = heads * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bf16* kb = k + (long long)b * Nk * C + hh * HD;
    const bf16* vb = v + (long long)b * Nk * C + hh * HD;
    const float inv_sqrt = sqrtf((float)HD);
    // ---- pass A: scores of all queries, one read of K
    {
        float qr[NQ][HD];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            const uint4* qp = reinterpret_cast<const uint4*>(q + ((long long)(q_bcast ? 0 : b) * Nq + (qi < Nq ? qi : 0)) * C + hh * HD);
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                const uint4 u = __ldg(qp + c8);
                const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                qr[qi][c8 * 8 + 0] = a.x; qr[qi][c8 * 8 + 1] = a.y; qr[qi][c8 * 8 + 2] = bb.x; qr[qi][c8 * 8 + 3] = bb.y;
                qr[qi][c8 * 8 + 4] = c.x; qr[qi][c8 * 8 + 5] = c.y; qr[qi][c8 * 8 + 6] = e.x; qr[qi][c8 * 8 + 7] = e.y;
            }
        }
        float lmax[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) lmax[qi] = -INFINITY;
        for (int j = threadIdx.x; j < Nk; j += 256) {
            const uint4* kp = reinterpret_cast<const uint4*>(kb + (long long)j * C);
            float kr[HD];
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                const uint4 u = __ldg(kp + c8);
                const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
                kr[c8 * 8 + 0] = a.x; kr[c8 * 8 + 1] = a.y; kr[c8 * 8 + 2] = bb.x; kr[c8 * 8 + 3] = bb.y;
                kr[c8 * 8 + 4] = c.x; kr[c8 * 8 + 5] = c.y; kr[c8 * 8 + 6] = e.x; kr[c8 * 8 + 7] = e.y;
            }
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                float d = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < HD / 8; ++c8) {
                    const float* qq = qr[qi] + c8 * 8;
                    const float* kk = kr + c8 * 8;
                    d += qq[0] * kk[0] + qq[1] * kk[1] + qq[2] * kk[2] + qq[3] * kk[3] + qq[4] * kk[4] + qq[5] * kk[5] + qq[6] * kk[6] +
                         qq[7] * kk[7];
                }
                d = bf16_round(bf16_round(d) / inv_sqrt);
                sc[(size_t)qi * Nk + j] = __float2bfloat16_rn(d);
                lmax[qi] = fmaxf(lmax[qi], d);
            }
        }
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            const float m = warp_max(lmax[qi]);
            if (lane == 0) red[warp * NQ + qi] = m;
        }
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        float m = red[threadIdx.x];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w * NQ + threadIdx.x]);
        stat[threadIdx.x] = m;
    }
    __syncthreads();
    // ---- pass B: sum of exp per query (shared memory only), then probabilities in place
    {
        float lsum[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) lsum[qi] = 0.f;
        for (int j = threadIdx.x; j < Nk; j += 256) {
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) lsum[qi] += __expf(__bfloat162float(sc[(size_t)qi * Nk + j]) - stat[qi]);
        }
        __syncthreads();   // every thread has read the maxima' partials out of red
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            const float t = warp_sum(lsum[qi]);
            if (lane == 0) red[warp * NQ + qi] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w * NQ + threadIdx.x];
        stat[NQ + threadIdx.x] = 1.f / t;
    }
    __syncthreads();
    // ---- pass C: probabilities (bf16, like the reference) times V, one read of V
    float acc[NQ][HD];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[qi][c] = 0.f;
    for (int j = threadIdx.x; j < Nk; j += 256) {
        const uint4* vp = reinterpret_cast<const uint4*>(vb + (long long)j * C);
        float vr[HD];
#pragma unroll
        for (int c8 = 0; c8 < HD / 8; ++c8) {
            const uint4 u = __ldg(vp + c8);
            const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
            vr[c8 * 8 + 0] = a.x; vr[c8 * 8 + 1] = a.y; vr[c8 * 8 + 2] = bb.x; vr[c8 * 8 + 3] = bb.y;
            vr[c8 * 8 + 4] = c.x; vr[c8 * 8 + 5] = c.y; vr[c8 * 8 + 6] = e.x; vr[c8 * 8 + 7] = e.y;
        }
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            const float pr = bf16_round(__expf(__bfloat162float(sc[(size_t)qi * Nk + j]) - stat[qi]) * stat[NQ + qi]);
#pragma unroll
            for (int c = 0; c < HD; ++c) acc[qi][c] += pr * vr[c];
        }
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            const float t = warp_sum(acc[qi][c]);
            if (lane == 0) red[warp * (NQ * HD) + qi * HD + c] = t;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < Nq * HD; i += 256) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w * (NQ * HD) + i];
        const int qi = i / HD, c = i % HD;
        out[((long long)b * Nq + qi) * C + hh * HD + c] = __float2bfloat16_rn(t);
    }
}

// Many queries, few keys (<=16): thread per (query, head), head fastest so q/out accesses are contiguous.
// Keys / values sit in shared memory as fp32 with one float of padding per head slice (8 heads at stride HD are 4 bank groups
// otherwise); the score array is indexed by unrolled loops only, so it stays in registers (a run-time key count used to put it
// in local memory: 0.15 ms per launch for 67 MB of traffic).
template <int HD>
__global__ void __launch_bounds__(256) attn_few_keys_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                            const bf16* __restrict__ v, bf16* __restrict__ out, int Nq,
                                                            int Nk, int heads) {
    const int b = blockIdx.y;
    const int C = heads * HD;
    constexpr int HP = HD + 1;
    extern __shared__ float kv_s[];  // k [Nk][heads][HD+1], v likewise, as fp32
    float* ks = kv_s;
    float* vs = kv_s + Nk * heads * HP;
    for (int i = threadIdx.x; i < Nk * C; i += blockDim.x) {
        const int o = (i / HD) * HP + (i % HD);
        ks[o] = __bfloat162float(k[(long long)b * Nk * C + i]);
        vs[o] = __bfloat162float(v[(long long)b * Nk * C + i]);
    }
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Nq * heads) return;
    const int hh = (int)(idx % heads);
    const long long qi = idx / heads;
    const uint4* qp = reinterpret_cast<const uint4*>(q + ((long long)b * Nq + qi) * C + hh * HD);
    float qr[HD];
#pragma unroll
    for (int c8 = 0; c8 < HD / 8; ++c8) {
        uint4 u = qp[c8];
        float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
        qr[c8 * 8 + 0] = a.x; qr[c8 * 8 + 1] = a.y; qr[c8 * 8 + 2] = bb.x; qr[c8 * 8 + 3] = bb.y;
        qr[c8 * 8 + 4] = c.x; qr[c8 * 8 + 5] = c.y; qr[c8 * 8 + 6] = e.x; qr[c8 * 8 + 7] = e.y;
    }
    const float inv_sqrt = sqrtf((float)HD);
    float sc[16];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        sc[j] = -INFINITY;
        if (j < Nk) {
            const float* kj = ks + (j * heads + hh) * HP;
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) d += qr[c] * kj[c];
            d = bf16_round(bf16_round(d) / inv_sqrt);
            sc[j] = d;
            mx = fmaxf(mx, d);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < Nk) {
            sc[j] = __expf(sc[j] - mx);
            sum += sc[j];
        }
    }
    const float inv = 1.f / sum;
    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < Nk) {
            const float pr = bf16_round(sc[j] * inv);
            const float* vj = vs + (j * heads + hh) * HP;
#pragma unroll
            for (int c = 0; c < HD; ++c) acc[c] += pr * vj[c];
        }
    }
    uint4* op = reinterpret_cast<uint4*>(out + ((long long)b * Nq + qi) * C + hh * HD);
#pragma unroll
    for (int c8 = 0; c8 < HD / 8; ++c8) {
        uint4 u;
        u.x = pack_bf16x2(acc[c8 * 8 + 0], acc[c8 * 8 + 1]);
        u.y = pack_bf16x2(acc[c8 * 8 + 2], acc[c8 * 8 + 3]);
        u.z = pack_bf16x2(acc[c8 * 8 + 4], acc[c8 * 8 + 5]);
        u.w = pack_bf16x2(acc[c8 * 8 + 6], acc[c8 * 8 + 7]);
        op[c8] = u;
    }
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_attention_bf16(ivlm_handle h, const ivlm_attn_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && a, "attention: null");
    IVLM_REQUIRE(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Sk > 0, "attention: empty problem");
    IVLM_REQUIRE(a->q_ts % 8 == 0 && a->k_ts % 8 == 0 && a->v_ts % 8 == 0 && a->q_hs % 8 == 0 && a->k_hs % 8 == 0 &&
                     a->v_hs % 8 == 0 && a->q_bs % 8 == 0 && a->k_bs % 8 == 0 && a->v_bs % 8 == 0,
                 "attention: q/k/v strides must be multiples of 8 elements (16-byte cp.async)");
    IVLM_REQUIRE(a->o_ts % 2 == 0 && a->o_hs % 2 == 0 && a->o_bs % 2 == 0, "attention: out strides must be even");
    const bool rel = a->rel_h != nullptr;
    IVLM_REQUIRE(!rel || (a->rel_w && a->kh * a->kw == a->Sk), "attention: rel-pos needs rel_w and kh*kw == Sk");
    IVLM_REQUIRE(!(rel && a->causal), "attention: rel-pos + causal is not a path the reference has");
    AttnParams p;
    p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.out = (bf16*)a->out;
    p.q_bs = a->q_bs; p.q_ts = a->q_ts; p.q_hs = a->q_hs;
    p.k_bs = a->k_bs; p.k_ts = a->k_ts; p.k_hs = a->k_hs;
    p.v_bs = a->v_bs; p.v_ts = a->v_ts; p.v_hs = a->v_hs;
    p.o_bs = a->o_bs; p.o_ts = a->o_ts; p.o_hs = a->o_hs;
    p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk;
    p.scale = a->scale;
    p.rel_h = a->rel_h; p.rel_w = a->rel_w; p.kh = a->kh; p.kw = a->kw;
    {   // CLIP (head_dim 64) and LLaMA prefill (128, causal) run on the tcgen05 kernel when the layout allows
        const int took = attention_tcgen05_try(h, a, stream);
        if (took < 0) return IVLM_ERR_CUDA;
        if (took == 1) {
            h->launches++;
            return IVLM_OK;
        }
    }
    if (a->D == 80) {
        return rel ? launch_fa<80, false, true>(h, p, stream) : launch_fa<80, false, false>(h, p, stream);
    } else if (a->D == 64) {
        IVLM_REQUIRE(!rel, "attention: rel-pos only instantiated for head_dim 80");
        return a->causal ? launch_fa<64, true, false>(h, p, stream) : launch_fa<64, false, false>(h, p, stream);
    } else if (a->D == 128) {
        IVLM_REQUIRE(!rel, "attention: rel-pos only instantiated for head_dim 80");
        return a->causal ? launch_fa<128, true, false>(h, p, stream) : launch_fa<128, false, false>(h, p, stream);
    }
    set_error("attention: head_dim %d not instantiated (64, 80, 128)", a->D);
    return IVLM_ERR_ARG;
}

extern "C" int ivlm_sam_relpos(ivlm_handle h, const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                               float* rel_h, float* rel_w, int32_t B, int32_t heads, int32_t Hq, int32_t Wq, int32_t hd,
                               void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && B > 0 && Hq > 0 && Wq > 0, "sam_relpos: empty");
    const size_t smem = (size_t)(Wq + Hq + 2 * Wq - 1) * (hd + 2) * 2;
    IVLM_CHECK_CUDA(cudaFuncSetAttribute(sam_relpos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(Hq, heads, B);
    sam_relpos_kernel<<<grid, 256, smem, stream>>>((const bf16*)qkv, (const bf16*)rel_pos_h, (const bf16*)rel_pos_w,
                                                   rel_h, rel_w, heads, Hq, Wq, hd);
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

extern "C" int ivlm_decode_attention_paged_bf16(ivlm_handle h, const void* q, const void* k_cache, const void* v_cache,
                                                const int32_t* block_table, const int32_t* seq_lens, void* out,
                                                int32_t B, int32_t H, int32_t hd, int32_t page_size, int32_t max_pages,
                                                float scale, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && B > 0 && H > 0, "decode_attention: empty");
    const size_t smem = (size_t)page_size * max_pages * sizeof(float);
    IVLM_REQUIRE(smem <= 160 * 1024, "decode_attention: max context %d too long for the score buffer",
                 page_size * max_pages);
    IVLM_REQUIRE(max_pages <= DEC_MAX_PAGES, "decode_attention: more than %d pages per sequence", DEC_MAX_PAGES);
    IVLM_REQUIRE(page_size == DEC_PAGE, "decode_attention: page_size %d not instantiated (%d)", page_size, DEC_PAGE);
    const float inv_scale = 1.0f / scale;  // reference divides by sqrt(hd)
    dim3 grid(H, B);
    // warps per CTA: every warp consumes whole pages.  8 is the default; 11 (two pages per warp at the bench's 22-page context
    // instead of three for six of the eight warps) and 16 measured slower, 17.1 us against 16.0 us per layer inside the decode
    // chain -- the extra warps lengthen the two block-wide reductions more than they shorten the page loops (option "dec_warps")
    const int warps = h->dec_warps == 11 ? 11 : (h->dec_warps == 16 ? 16 : 8);
#define IVLM_DEC(HD_, W_, PF_)                                                                                                   \
    do {                                                                                                                         \
        IVLM_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_paged_kernel<HD_, W_, PF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                                        \
        IVLM_CHECK_CUDA(launch_k(h, decode_attn_paged_kernel<HD_, W_, PF_>, grid, dim3(W_ * 32), smem, stream, (const bf16*)q,    \
                                 (const bf16*)k_cache, (const bf16*)v_cache, (const int*)block_table, (const int*)seq_lens,      \
                                 (bf16*)out, (int)H, (int)max_pages, inv_scale));                                                \
    } while (0)
    const bool pf = h->dec_prefetch != 0;   // option "dec_prefetch" (default off, measured slower): L2 requests for the K / V working set up front
    if (hd == 128) {
        if (warps == 8) { if (pf) IVLM_DEC(128, 8, true); else IVLM_DEC(128, 8, false); }
        else if (warps == 16) IVLM_DEC(128, 16, false);
        else IVLM_DEC(128, 11, false);
    } else if (hd == 64) {
        if (pf) IVLM_DEC(64, 8, true); else IVLM_DEC(64, 8, false);
    } else {
        set_error("decode_attention: head_dim %d not instantiated (64, 128)", hd);
        return IVLM_ERR_ARG;
    }
#undef IVLM_DEC
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

extern "C" int ivlm_attn_small_bf16(ivlm_handle h, const void* q, const void* k, const void* v, void* out, int32_t B,
                                    int32_t q_bcast, int32_t Nq, int32_t Nk, int32_t heads, int32_t hd, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && B > 0 && Nq > 0 && Nk > 0, "attn_small: empty");
    IVLM_REQUIRE(hd == 16 || hd == 32, "attn_small: head_dim %d not instantiated (16, 32)", hd);
    if (Nk <= 16) {
        IVLM_REQUIRE(!q_bcast, "attn_small: broadcast q unsupported in the few-keys form");
        const size_t smem = (size_t)2 * Nk * heads * (hd + 1) * sizeof(float);
        dim3 grid((unsigned)(((long long)Nq * heads + 255) / 256), B);
        if (hd == 16)
            attn_few_keys_kernel<16><<<grid, 256, smem, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                                  (bf16*)out, Nq, Nk, heads);
        else
            attn_few_keys_kernel<32><<<grid, 256, smem, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                                  (bf16*)out, Nq, Nk, heads);
    } else {
        IVLM_REQUIRE(Nq <= 64, "attn_small: needs Nq <= 64 or Nk <= 16 (got Nq=%d Nk=%d)", Nq, Nk);
        dim3 grid(heads, B);
        // all queries in one sweep when the [NQ][Nk] bf16 score tile fits in shared memory (the SAM decoder: 9 tokens x 4096 keys)
        const int nq_t = Nq <= 10 ? 10 : 16;
        const size_t smem_all = (size_t)nq_t * Nk * 2 + (size_t)(8 * nq_t * 16 + 2 * nq_t) * sizeof(float);
        if (hd == 16 && Nq <= 16 && smem_all <= 200 * 1024 && h->attn_small_variant == 0) {
            if (!(h->attr_done & (1ull << 21))) {
                IVLM_CHECK_CUDA(cudaFuncSetAttribute(attn_few_queries_all_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                IVLM_CHECK_CUDA(cudaFuncSetAttribute(attn_few_queries_all_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                h->attr_done |= 1ull << 21;
            }
            if (nq_t == 10)
                attn_few_queries_all_kernel<10><<<grid, 256, smem_all, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out,
                                                                               q_bcast, Nq, Nk, heads);
            else
                attn_few_queries_all_kernel<16><<<grid, 256, smem_all, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out,
                                                                               q_bcast, Nq, Nk, heads);
        } else if (hd == 16)
            attn_few_queries_kernel<16><<<grid, 256, 0, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                                  (bf16*)out, q_bcast, Nq, Nk, heads);
        else
            attn_few_queries_kernel<32><<<grid, 256, 0, stream>>>((const bf16*)q, (const bf16*)k, (const bf16*)v,
                                                                  (bf16*)out, q_bcast, Nq, Nk, heads);
    }
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}
