// RoPE + paged KV-cache store (LLaMA), and the fused tail of the SAM mask decoder
// (second ConvTranspose2d + GELU + hypernetwork dot product).
#include "common.cuh"
#include "runtime.h"

namespace ivlm {

// HF apply_rotary_pos_emb (transformers 4.31): q_embed = q*cos + rotate_half(q)*sin with bf16 cos/sin tables;
// every product and the sum are bf16 ops in the reference, so each is rounded here.
// One CTA per token; thread i handles (head, pair) items.
__global__ void rope_kv_store_kernel(const bf16* __restrict__ qkv, const int* __restrict__ positions,
                                     const int* __restrict__ slot_map, const bf16* __restrict__ cos_t,
                                     const bf16* __restrict__ sin_t, bf16* __restrict__ q_out, bf16* __restrict__ k_out,
                                     bf16* __restrict__ v_out, bf16* __restrict__ k_cache, bf16* __restrict__ v_cache,
                                     int H, int hd, int page, int paired) {
    pdl_wait_then_launch();
    // grid (tokens, chunks): every thread owns one (head, rotary pair) and one 16-byte piece of the V row, so a
    // decode step (8 tokens) spreads over ~100 CTAs instead of serialising ten dependent loads per thread in 8
    const long long tkn = blockIdx.x;
    const int pos = positions[tkn];
    const long long slot = slot_map ? slot_map[tkn] : -1;
    // paged cache layout [pages, H, page, hd]: the `page` tokens of one head are contiguous, so a decode CTA streams
    // page*hd*2-byte runs instead of hd*2-byte pieces strided by H*hd
    const long long pg = slot >= 0 ? slot / page : 0, off = slot >= 0 ? slot % page : 0;
    auto cidx = [&](int hh, int j) { return ((pg * H + hh) * page + off) * hd + j; };
    const int half = hd >> 1;
    const int D = H * hd;
    const bf16* qr = qkv + tkn * 3LL * D;
    const bf16* kr = qr + D;
    const bf16* vr = kr + D;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < H * half; i += gridDim.y * blockDim.x) {
        const int hh = i / half, j = i % half;
        const float c = __bfloat162float(cos_t[(long long)pos * hd + j]);
        const float s = __bfloat162float(sin_t[(long long)pos * hd + j]);
        const int i1 = hh * hd + j, i2 = i1 + half;
        // `paired` input columns (the row order ivlm_decode_linear's ROPE_KV epilogue needs, used for q and k by every path so
        // that the weights exist once): feature j of a head sits at (j / 8) * 16 + j % 8, its partner j + half 8 columns later
        const int s1 = paired ? hh * hd + (j >> 3) * 16 + (j & 7) : i1, s2 = paired ? s1 + 8 : i2;
        {
            const float x1 = __bfloat162float(qr[s1]), x2 = __bfloat162float(qr[s2]);
            q_out[tkn * D + i1] = __float2bfloat16_rn(bf16_round(x1 * c) + bf16_round(-x2 * s));
            q_out[tkn * D + i2] = __float2bfloat16_rn(bf16_round(x2 * c) + bf16_round(x1 * s));
        }
        {
            const float x1 = __bfloat162float(kr[s1]), x2 = __bfloat162float(kr[s2]);
            const bf16 o1 = __float2bfloat16_rn(bf16_round(x1 * c) + bf16_round(-x2 * s));
            const bf16 o2 = __float2bfloat16_rn(bf16_round(x2 * c) + bf16_round(x1 * s));
            if (k_out) { k_out[tkn * D + i1] = o1; k_out[tkn * D + i2] = o2; }
            if (slot >= 0) { k_cache[cidx(hh, j)] = o1; k_cache[cidx(hh, j + half)] = o2; }
        }
    }
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < (D >> 3); i += gridDim.y * blockDim.x) {
        const uint4 u = reinterpret_cast<const uint4*>(vr)[i];
        if (v_out) reinterpret_cast<uint4*>(v_out + tkn * D)[i] = u;
        if (slot >= 0) *reinterpret_cast<uint4*>(v_cache + cidx((i * 8) / hd, (i * 8) % hd)) = u;
    }
}

// up1: [Bv, G*G tokens, 4 (dy1*2+dx1), 64] bf16 (LayerNorm2d+GELU already applied).
// For each first-stage pixel: z[p2][c] = gelu(bf16(sum_k u[k] W2[p2][c][k] + b2[c])), mask = bf16(sum_c hyper[c] z[p2][c]).
// Warp = 32 consecutive first-stage pixels with one p2 => shared-memory weight reads are broadcasts.
__global__ void __launch_bounds__(128) upscale_hyper_dot_kernel(const bf16* __restrict__ up1, const bf16* __restrict__ w2,
                                                                const bf16* __restrict__ b2,
                                                                const bf16* __restrict__ hyper, float* __restrict__ lowres,
                                                                int G) {
    __shared__ float w_s[4 * 32 * 64];
    __shared__ float b_s[32], h_s[32];
    const int bv = blockIdx.y;
    for (int i = threadIdx.x; i < 4 * 32 * 64; i += blockDim.x) w_s[i] = __bfloat162float(w2[i]);
    if (threadIdx.x < 32) {
        b_s[threadIdx.x] = __bfloat162float(b2[threadIdx.x]);
        h_s[threadIdx.x] = __bfloat162float(hyper[bv * 32 + threadIdx.x]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p2 = warp;  // 4 warps <-> 4 second-stage sub-pixels
    const long long vec = (long long)blockIdx.x * 32 + lane;  // first-stage pixel index in [0, G*G*4)
    const long long nvec = (long long)G * G * 4;
    if (vec >= nvec) return;
    const uint4* up = reinterpret_cast<const uint4*>(up1 + ((long long)bv * nvec + vec) * 64);
    float u[64];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        uint4 q = up[c8];
        float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
        u[c8 * 8 + 0] = a.x; u[c8 * 8 + 1] = a.y; u[c8 * 8 + 2] = b.x; u[c8 * 8 + 3] = b.y;
        u[c8 * 8 + 4] = c.x; u[c8 * 8 + 5] = c.y; u[c8 * 8 + 6] = d.x; u[c8 * 8 + 7] = d.y;
    }
    float m = 0.f;
    const float* wp = w_s + p2 * 32 * 64;
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
        float z = 0.f;
#pragma unroll
        for (int k = 0; k < 64; ++k) z += u[k] * wp[c * 64 + k];
        z = bf16_round(z + b_s[c]);
        z = bf16_round(apply_act(z, ACT_GELU));
        m += h_s[c] * z;
    }
    const long long tok = vec >> 2;
    const int p1 = (int)(vec & 3);
    const int y = (int)(tok / G), x = (int)(tok % G);
    const int Y = (y * 2 + (p1 >> 1)) * 2 + (p2 >> 1);
    const int X = (x * 2 + (p1 & 1)) * 2 + (p2 & 1);
    const int LR = G * 4;
    lowres[((long long)bv * LR + Y) * LR + X] = bf16_round(m);
}

// The same computation as a tensor-core GEMM [pixels, 64] x [64, 4*32] with the bias / GELU / hypernetwork dot as its epilogue
// (mma.sync m16n8k16: K = 64 and 2 KB of operands per 16-pixel tile leave nothing for a TMA / tcgen05 pipeline to hide; the
// launch reads up1 once -- 67 MB at 32 views -- and writes 8 MB).  The scalar kernel above spends 2048 shared-memory-fed FMAs per
// pixel (0.42-0.54 ms per launch); it stays as the option-selected A/B form.  The k index is permuted consistently in both
// operands (physical k = 16 t + 4 ks + e for fragment slot e of k-step ks) so that a thread's A fragments are the 32 contiguous
// bytes it loads with two 16-byte requests and its B fragments one 16-byte shared-memory read per two k-steps.
constexpr int UH_PITCH = 72;   // bf16 per weight row in shared memory: 36 words -> conflict-free 16-byte fragment reads
__global__ void __launch_bounds__(128) upscale_hyper_dot_mma_kernel(const bf16* __restrict__ up1, const bf16* __restrict__ w2,
                                                                    const bf16* __restrict__ b2, const bf16* __restrict__ hyper,
                                                                    float* __restrict__ lowres, int G, int tiles_per_cta) {
    __shared__ __align__(16) bf16 w_s[128 * UH_PITCH];
    __shared__ float b_s[32], h_s[32];
    const int bv = blockIdx.y;
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {   // 128 rows x 8 chunks of 16 bytes
        const int n = i >> 3, c8 = i & 7;
        *reinterpret_cast<uint4*>(w_s + n * UH_PITCH + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(w2) + i);
    }
    if (threadIdx.x < 32) {
        b_s[threadIdx.x] = __bfloat162float(b2[threadIdx.x]);
        h_s[threadIdx.x] = __bfloat162float(hyper[bv * 32 + threadIdx.x]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const long long nvec = (long long)G * G * 4;
    const int LR = G * 4;
    const long long tile0 = (long long)blockIdx.x * tiles_per_cta;
    for (int it = warp; it < tiles_per_cta; it += 4) {
        const long long r0 = (tile0 + it) * 16 + g, r1 = r0 + 8;
        if ((tile0 + it) * 16 >= nvec) break;
        uint4 a0lo = make_uint4(0, 0, 0, 0), a0hi = a0lo, a1lo = a0lo, a1hi = a0lo;
        if (r0 < nvec) {
            const uint4* p0 = reinterpret_cast<const uint4*>(up1 + ((long long)bv * nvec + r0) * 64 + t * 16);
            a0lo = __ldg(p0); a0hi = __ldg(p0 + 1);
        }
        if (r1 < nvec) {
            const uint4* p1 = reinterpret_cast<const uint4*>(up1 + ((long long)bv * nvec + r1) * 64 + t * 16);
            a1lo = __ldg(p1); a1hi = __ldg(p1 + 1);
        }
        // fragment registers of k-step ks: {row g: slots 0-1, row g+8: slots 0-1, row g: slots 2-3, row g+8: slots 2-3}
        const uint32_t af[4][4] = {{a0lo.x, a1lo.x, a0lo.y, a1lo.y}, {a0lo.z, a1lo.z, a0lo.w, a1lo.w},
                                   {a0hi.x, a1hi.x, a0hi.y, a1hi.y}, {a0hi.z, a1hi.z, a0hi.w, a1hi.w}};
        float m0[4] = {0.f, 0.f, 0.f, 0.f}, m1[4] = {0.f, 0.f, 0.f, 0.f};   // per second-stage sub-pixel p2: rows g, g+8
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint4* bp = reinterpret_cast<const uint4*>(w_s + (j * 8 + g) * UH_PITCH + t * 16);
            const uint4 blo = bp[0], bhi = bp[1];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(acc, af[0], blo.x, blo.y);
            mma_bf16_16816(acc, af[1], blo.z, blo.w);
            mma_bf16_16816(acc, af[2], bhi.x, bhi.y);
            mma_bf16_16816(acc, af[3], bhi.z, bhi.w);
            const int c = (j & 3) * 8 + 2 * t;
            const float bz0 = b_s[c], bz1 = b_s[c + 1], h0 = h_s[c], h1 = h_s[c + 1];
            const float z00 = bf16_round(apply_act(bf16_round(acc[0] + bz0), ACT_GELU));
            const float z01 = bf16_round(apply_act(bf16_round(acc[1] + bz1), ACT_GELU));
            const float z10 = bf16_round(apply_act(bf16_round(acc[2] + bz0), ACT_GELU));
            const float z11 = bf16_round(apply_act(bf16_round(acc[3] + bz1), ACT_GELU));
            m0[j >> 2] += h0 * z00 + h1 * z01;
            m1[j >> 2] += h0 * z10 + h1 * z11;
        }
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
            m0[p2] += __shfl_xor_sync(0xffffffffu, m0[p2], 1);
            m0[p2] += __shfl_xor_sync(0xffffffffu, m0[p2], 2);
            m1[p2] += __shfl_xor_sync(0xffffffffu, m1[p2], 1);
            m1[p2] += __shfl_xor_sync(0xffffffffu, m1[p2], 2);
        }
        // lane t stores sub-pixel p2 = t of both rows
        const float v0 = t == 0 ? m0[0] : t == 1 ? m0[1] : t == 2 ? m0[2] : m0[3];
        const float v1 = t == 0 ? m1[0] : t == 1 ? m1[1] : t == 2 ? m1[2] : m1[3];
        const int p2 = t;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const long long vec = half ? r1 : r0;
            if (vec >= nvec) continue;
            const long long tok = vec >> 2;
            const int p1 = (int)(vec & 3);
            const int y = (int)(tok / G), x = (int)(tok % G);
            const int Y = (y * 2 + (p1 >> 1)) * 2 + (p2 >> 1);
            const int X = (x * 2 + (p1 & 1)) * 2 + (p2 & 1);
            lowres[((long long)bv * LR + Y) * LR + X] = bf16_round(half ? v1 : v0);
        }
    }
}

// Device-side bookkeeping of the greedy / scripted decode loop, so that a decode step is a pure graph replay.
// state[0] = number of tokens fed so far (advanced here), state[1] = the step the rest of this replay works on.
__global__ void decode_prepare_kernel(int* __restrict__ state, int S, const int* __restrict__ S_rows, const int* __restrict__ scripted, int G,
                                      const int* __restrict__ next, int* __restrict__ done, int* __restrict__ out_tokens,
                                      int* __restrict__ tok, int* __restrict__ pos, int* __restrict__ slot,
                                      int* __restrict__ seq_lens, const int* __restrict__ slot_base, int eos, int pad, int B) {
    pdl_wait_then_launch();
    const int step = state[0];
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int t = scripted != nullptr ? scripted[b * G + step] : next[b];
        if (done[b]) t = pad;
        out_tokens[b * G + step] = t;
        if (t == eos) done[b] = 1;
        const int p = (S_rows != nullptr ? S_rows[b] : S) + step;   // per-sample prompt length (right-padded batches)
        tok[b] = t;
        pos[b] = p;
        slot[b] = slot_base[b] + p;
        seq_lens[b] = p + 1;
    }
    if (threadIdx.x == 0) {
        state[1] = step;
        state[0] = step + 1;
    }
}
// hidden[b, S + step, :] = hid_step[b, :]
__global__ void decode_finish_kernel(const int* __restrict__ state, int S, const int* __restrict__ S_rows,
                                     const bf16* __restrict__ hid_step, bf16* __restrict__ hidden, int D, int max_len) {
    pdl_wait_then_launch();
    const int b = blockIdx.x, p = (S_rows != nullptr ? S_rows[b] : S) + state[1];
    const uint4* src = reinterpret_cast<const uint4*>(hid_step + (long long)b * D);
    uint4* dst = reinterpret_cast<uint4*>(hidden + ((long long)b * max_len + p) * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src[i];
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_decode_prepare(ivlm_handle h, int32_t* state, int32_t S, const int32_t* S_rows, const int32_t* scripted, int32_t G,
                                   const int32_t* next, int32_t* done, int32_t* out_tokens, int32_t* tok, int32_t* pos,
                                   int32_t* slot, int32_t* seq_lens, const int32_t* slot_base, int32_t eos, int32_t pad,
                                   int32_t B, void* stream) {
    IVLM_REQUIRE(h && state && next && done && out_tokens && tok && pos && slot && seq_lens && slot_base && B > 0 && G > 0,
                 "decode_prepare: bad arguments");
    IVLM_CHECK_CUDA(launch_k(h, decode_prepare_kernel, dim3(1), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), (int*)state,
                             (int)S, (const int*)S_rows, (const int*)scripted, (int)G, (const int*)next, (int*)done, (int*)out_tokens, (int*)tok,
                             (int*)pos, (int*)slot, (int*)seq_lens, (const int*)slot_base, (int)eos, (int)pad, (int)B));
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

extern "C" int ivlm_decode_finish(ivlm_handle h, const int32_t* state, int32_t S, const int32_t* S_rows, const void* hid_step, void* hidden,
                                  int32_t B, int32_t D, int32_t max_len, void* stream) {
    IVLM_REQUIRE(h && state && hid_step && hidden && B > 0 && D % 8 == 0, "decode_finish: bad arguments");
    IVLM_CHECK_CUDA(launch_k(h, decode_finish_kernel, dim3(B), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                             (const int*)state, (int)S, (const int*)S_rows, (const bf16*)hid_step, (bf16*)hidden, (int)D, (int)max_len));
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

extern "C" int ivlm_rope_kv_store_bf16(ivlm_handle h, const void* qkv, const int32_t* positions, const int32_t* slot_map,
                                       const void* cos_t, const void* sin_t, void* q_out, void* k_out, void* v_out,
                                       void* k_cache, void* v_cache, int32_t T, int32_t H, int32_t hd, int32_t page_size,
                                       int32_t paired, void* stream) {
    IVLM_REQUIRE(h && T > 0 && (H * hd) % 8 == 0 && hd % 8 == 0, "rope: bad shape");
    IVLM_REQUIRE(!paired || hd % 16 == 0, "rope: the paired column layout needs head_dim %% 16 == 0");
    IVLM_REQUIRE(slot_map == nullptr || page_size > 0, "rope: page_size must be positive when a cache is written");
    IVLM_REQUIRE(slot_map == nullptr || (k_cache && v_cache), "rope: slot_map given without caches");
    const int items = H * (hd / 2);
    int chunks = (items + 255) / 256;
    if ((long long)T * chunks > 4LL * h->num_sms) chunks = (int)((4LL * h->num_sms + T - 1) / T);  // prefill: fewer, looping CTAs
    if (chunks < 1) chunks = 1;
    IVLM_CHECK_CUDA(launch_k(h, rope_kv_store_kernel, dim3(T, chunks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                             (const bf16*)qkv, (const int*)positions, (const int*)slot_map, (const bf16*)cos_t, (const bf16*)sin_t,
                             (bf16*)q_out, (bf16*)k_out, (bf16*)v_out, (bf16*)k_cache, (bf16*)v_cache, (int)H, (int)hd,
                             (int)(page_size > 0 ? page_size : 1), (int)(paired ? 1 : 0)));
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}

extern "C" int ivlm_upscale_hyper_dot(ivlm_handle h, const void* up1, const void* w2, const void* b2, const void* hyper,
                                      float* lowres, int32_t Bv, int32_t grid_, void* stream) {
    IVLM_REQUIRE(h && Bv > 0 && grid_ > 0, "upscale_hyper_dot: empty");
    const long long nvec = (long long)grid_ * grid_ * 4;
    if (h->attn_small_variant == 0 && (reinterpret_cast<uintptr_t>(up1) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0) {
        const long long tiles = (nvec + 15) / 16;
        const int per_cta = tiles >= 64 ? 16 : 4;
        dim3 grid((unsigned)((tiles + per_cta - 1) / per_cta), Bv);
        upscale_hyper_dot_mma_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            (const bf16*)up1, (const bf16*)w2, (const bf16*)b2, (const bf16*)hyper, lowres, grid_, per_cta);
    } else {
        dim3 grid((unsigned)((nvec + 31) / 32), Bv);
        upscale_hyper_dot_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            (const bf16*)up1, (const bf16*)w2, (const bf16*)b2, (const bf16*)hyper, lowres, grid_);
    }
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}
