// out[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual for SMALL token counts M <= 64 (LLaMA decode steps, SAM decoder
// tokens, lm_head, [SEG] projection).  These launches stream every weight once and do 2*M FLOP per weight element: they are
// HBM-bound, so the kernel is built around the weight stream, not around a tensor-core tile:
//   * one CTA per 16 weight rows (N/16 CTAs: 320 ... 2000 on this path -> every SM streams, no split-K exchange);
//   * the k range of those rows is dealt to the CTA's warps in chunks of UNROLL adjacent 32-element steps; each lane issues
//     16-byte loads (8 consecutive k of one row), UNROLL of them in flight, so a warp reads 256 contiguous bytes of every
//     row per iteration and a round of the CTA's warps 1-2 KB;
//   * the arithmetic rides on mma.sync m16n8k16 (weights = the 16-row operand, 8 tokens = the n operand) with a k-slot
//     permutation that makes a lane's 16-byte load exactly its fragment (dot products do not care about k order as long as
//     both operands use the same one);
//   * warps' partial tiles are summed in warp order in shared memory (deterministic), then bias / activation / residual
//     with the same bf16 rounding points as the tcgen05 epilogue, and a transposed store out[token, row].
#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int GV_ROWS = 16, GV_UNROLL = 4, GV_MAX_WARPS = 16;

struct GemvParams {
    const bf16* a;   // tokens [M, K]
    long long lda;
    const bf16* w;   // weights [N, K]
    long long ldw;
    void* out;
    long long ldo;
    const bf16* bias;
    const bf16* res;
    long long ldr;
    int M, N, K;
    int act, out_f32, round_steps;
};

// NT: token groups of 8.  ROWS: weight rows per CTA, 16 (both halves of the m16 operand) or 8 (upper half zero) -- the
// latter doubles the CTA count for narrow layers (N = 5120: 640 CTAs instead of 320) so that the per-SM share of the
// stream stays within what one SM can pull (~60 GB/s) even on the SMs that receive one CTA more than the average.
template <int NT, int ROWS>
__global__ void __launch_bounds__(GV_MAX_WARPS * 32) gemv_small_m_kernel(const GemvParams p) {
    extern __shared__ float part[];  // [warps][GV_ROWS][8*NT + 1]
    constexpr int PLD = 8 * NT + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;  // the k-steps of this CTA's 16 rows are dealt round-robin to its warps
    const int g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.x * ROWS;
    const int r_lo = min(row0 + g, p.N - 1), r_hi = min(row0 + g + 8, p.N - 1);
    const bf16* w_lo = p.w + (long long)r_lo * p.ldw + 8 * t;
    const bf16* w_hi = p.w + (long long)r_hi * p.ldw + 8 * t;
    const int steps = p.K >> 5;  // 32 k per step
    const int stride = nwarps * GV_UNROLL;

    // a warp iteration covers GV_UNROLL ADJACENT k-steps (GV_UNROLL * 64 contiguous bytes of every row); the chunks of
    // one round are dealt to consecutive warps, so a round reads nwarps * GV_UNROLL * 64 contiguous bytes per row.
    // The FIRST round is issued before griddepcontrol.wait -- weights are static, so under programmatic dependent launch
    // they stream while the predecessor kernel is still running.  (Double buffering the weight registers costs 30
    // registers and the third resident CTA per SM: 320 CTAs would no longer fit one wave.)
    uint4 wl[GV_UNROLL], wh[GV_UNROLL];
    auto load_round = [&](uint4(&l)[GV_UNROLL], uint4(&h)[GV_UNROLL], int s0) {
#pragma unroll
        for (int u = 0; u < GV_UNROLL; ++u) {
            const int s = s0 + u;
            if (s < steps) {
                l[u] = __ldcs(reinterpret_cast<const uint4*>(w_lo + (s << 5)));  // streamed once: evict-first
                h[u] = ROWS == 16 ? __ldcs(reinterpret_cast<const uint4*>(w_hi + (s << 5))) : make_uint4(0, 0, 0, 0);
            }
        }
    };
    load_round(wl, wh, warp * GV_UNROLL);
    pdl_launch();
    pdl_wait();

    const bf16* x_row[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) x_row[nt] = p.a + (long long)min(nt * 8 + g, p.M - 1) * p.lda + 8 * t;

    float c[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;

    for (int s0 = warp * GV_UNROLL; s0 < steps; s0 += stride) {
#pragma unroll
        for (int u = 0; u < GV_UNROLL; ++u) {
            const int s = s0 + u;
            if (s < steps) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint4 xb = *reinterpret_cast<const uint4*>(x_row[nt] + (s << 5));
                    const uint32_t a1[4] = {wl[u].x, wh[u].x, wl[u].y, wh[u].y};
                    const uint32_t a2[4] = {wl[u].z, wh[u].z, wl[u].w, wh[u].w};
                    mma_bf16_16816(c[nt], a1, xb.x, xb.y);
                    mma_bf16_16816(c[nt], a2, xb.z, xb.w);
                }
            }
        }
        load_round(wl, wh, s0 + stride);
    }
    // c[nt][0,1] = (row g, tokens nt*8 + 2t, 2t+1); c[nt][2,3] = (row g+8, same tokens)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        float* pw = part + (warp * GV_ROWS) * PLD;
        pw[g * PLD + nt * 8 + 2 * t] = c[nt][0];
        pw[g * PLD + nt * 8 + 2 * t + 1] = c[nt][1];
        if (ROWS == 16) {
            pw[(g + 8) * PLD + nt * 8 + 2 * t] = c[nt][2];
            pw[(g + 8) * PLD + nt * 8 + 2 * t + 1] = c[nt][3];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ROWS * 8 * NT; i += blockDim.x) {
        const int rl = i % ROWS, tok = i / ROWS;  // consecutive threads -> consecutive rows of one token
        const int row = row0 + rl;
        if (row >= p.N || tok >= p.M) continue;
        float x = 0.f;
        for (int w = 0; w < nwarps; ++w) x += part[(w * GV_ROWS + rl) * PLD + tok];  // fixed order: deterministic
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
}

int launch_gemv_small_m(ivlm_ctx* h, const ivlm_gemm_args* a, cudaStream_t stream) {
    GemvParams p;
    p.a = reinterpret_cast<const bf16*>(a->a); p.lda = a->lda;
    p.w = reinterpret_cast<const bf16*>(a->w); p.ldw = a->ldw;
    p.out = a->out; p.ldo = a->ldo;
    p.bias = reinterpret_cast<const bf16*>(a->bias);
    p.res = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.act = a->act;
    p.out_f32 = a->out_dtype == IVLM_F32;
    p.round_steps = (a->out_dtype == IVLM_BF16 && !a->no_round) ? 1 : 0;
    const int rows = (a->N <= h->gv_rows8_max_n) ? 8 : 16;
    const int ctas = (a->N + rows - 1) / rows;
    const int nt = (a->M + 7) / 8;
    const int NT = nt <= 1 ? 1 : (nt <= 2 ? 2 : (nt <= 4 ? 4 : 8));
    // enough warps per SM to cover the HBM latency (~24), within the k-steps available and 48 KB of shared memory
    int warps = (24 * h->num_sms + ctas - 1) / ctas;
    warps = warps < 4 ? 4 : (warps > 8 ? 8 : warps);
    if (h->gv_warps > 0) warps = h->gv_warps;
    while (warps > 1 && (size_t)warps * GV_ROWS * (8 * NT + 1) * sizeof(float) > 48 * 1024) warps >>= 1;
    while (warps > 1 && warps * GV_UNROLL > (a->K >> 5)) warps >>= 1;
    const size_t smem = (size_t)warps * GV_ROWS * (8 * NT + 1) * sizeof(float);
    const dim3 grid(ctas), block(warps * 32);
    cudaError_t e;
#define IVLM_GV(NT_)                                                                                      \
    e = rows == 16 ? launch_k(h, gemv_small_m_kernel<NT_, 16>, grid, block, smem, stream, p)              \
                   : launch_k(h, gemv_small_m_kernel<NT_, 8>, grid, block, smem, stream, p)
    switch (NT) {
        case 1: IVLM_GV(1); break;
        case 2: IVLM_GV(2); break;
        case 4: IVLM_GV(4); break;
        default: IVLM_GV(8); break;
    }
#undef IVLM_GV
    if (e != cudaSuccess) {
        set_error("gemv_small_m launch failed: %s", cudaGetErrorString(e));
        return IVLM_ERR_CUDA;
    }
    h->launches++;
    return IVLM_OK;
}

}  // namespace ivlm
