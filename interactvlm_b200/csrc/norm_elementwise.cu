// HBM-bound row-wise and elementwise kernels: LayerNorm / RMSNorm (warp per row, 16-byte loads, warp-shuffle
// reductions, fp32 statistics), broadcast add, SwiGLU gate, split-K finalize, casts, im2col lowering,
// embedding gather / image splice, greedy argmax, bilinear resize, camera gate.
#include <algorithm>

#include "common.cuh"
#include "runtime.h"

namespace ivlm {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per output row; D multiple of 8. x is re-read from L1 for the second/third pass.
__global__ void layernorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const bf16* __restrict__ gamma,
                                 const bf16* __restrict__ beta, long long out_rows, int D, float eps,
                                 const int* __restrict__ row_map, int act) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= out_rows) return;
    const int lane = threadIdx.x & 31;
    long long src = row;
    if (row_map != nullptr) src = row_map[row];
    bf16* yr = y + row * D;
    const int nvec = D >> 3;
    if (src < 0) {  // window padding rows are zeros AFTER the norm (image_encoder.py:179-183)
        for (int i = lane; i < nvec; i += 32) reinterpret_cast<uint4*>(yr)[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint4* xr = reinterpret_cast<const uint4*>(x + src * D);
    float s = 0.f;
    for (int i = lane; i < nvec; i += 32) {
        uint4 q = xr[i];
        float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
        s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
    for (int i = lane; i < nvec; i += 32) {
        uint4 q = xr[i];
        float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
        float t;
        t = a.x - mean; ss += t * t; t = a.y - mean; ss += t * t;
        t = b.x - mean; ss += t * t; t = b.y - mean; ss += t * t;
        t = c.x - mean; ss += t * t; t = c.y - mean; ss += t * t;
        t = d.x - mean; ss += t * t; t = d.y - mean; ss += t * t;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    const uint4* b4 = reinterpret_cast<const uint4*>(beta);
    for (int i = lane; i < nvec; i += 32) {
        uint4 q = xr[i], g = g4[i], b = b4[i];
        uint32_t xi[4] = {q.x, q.y, q.z, q.w}, gi[4] = {g.x, g.y, g.z, g.w}, bi[4] = {b.x, b.y, b.z, b.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]), bv = unpack_bf16x2(bi[j]);
            float r0 = (xv.x - mean) * rstd * gv.x + bv.x;
            float r1 = (xv.y - mean) * rstd * gv.y + bv.y;
            if (act != ACT_NONE) {
                r0 = apply_act(bf16_round(r0), act);
                r1 = apply_act(bf16_round(r1), act);
            }
            o[j] = pack_bf16x2(r0, r1);
        }
        reinterpret_cast<uint4*>(yr)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Register-resident variant for D = 256 * VPL (SAM 1280, CLIP 1024, decoder 256): the row is read from HBM exactly once
// (VPL 16-byte vectors per lane), statistics and the normalised output come from registers.
template <int VPL>
__global__ void __launch_bounds__(256, 6) layernorm_reg_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                            const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                                            long long out_rows, float eps, const int* __restrict__ row_map,
                                                            int act) {
    constexpr int D = VPL * 256;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= out_rows) return;
    const int lane = threadIdx.x & 31;
    long long src = row;
    if (row_map != nullptr) src = row_map[row];
    uint4* yr = reinterpret_cast<uint4*>(y + row * D);
    if (src < 0) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) yr[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint4* xr = reinterpret_cast<const uint4*>(x + src * D);
    // the row stays PACKED in registers (VPL x 4 instead of VPL x 8 words) and is unpacked in each of the three passes: the
    // kernel is HBM-bound and 40 instead of 70 registers double the resident warps (loads in flight per SM)
    uint4 q[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) q[k] = xr[lane + 32 * k];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const float2 a = unpack_bf16x2(q[k].x), b = unpack_bf16x2(q[k].y), c = unpack_bf16x2(q[k].z), d = unpack_bf16x2(q[k].w);
        s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
    }
    const float mean = warp_sum(s) / (float)D;
    // opaque to the optimiser: without it the unpacked floats of the first pass are kept live for the next two
#define IVLM_LN_REPACK()                                                                                   \
    _Pragma("unroll") for (int k = 0; k < VPL; ++k)                                                        \
        asm volatile("" : "+r"(q[k].x), "+r"(q[k].y), "+r"(q[k].z), "+r"(q[k].w))
    IVLM_LN_REPACK();
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const uint32_t w4[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t2 = unpack_bf16x2(w4[j]);
            const float t0 = t2.x - mean, t1 = t2.y - mean;
            ss += t0 * t0;
            ss += t1 * t1;
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
    IVLM_LN_REPACK();
#undef IVLM_LN_REPACK
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    const uint4* b4 = reinterpret_cast<const uint4*>(beta);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const uint4 g = g4[lane + 32 * k], b = b4[lane + 32 * k];
        const uint32_t xi[4] = {q[k].x, q[k].y, q[k].z, q[k].w}, gi[4] = {g.x, g.y, g.z, g.w}, bi[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]), bv = unpack_bf16x2(bi[j]);
            float r0 = (xv.x - mean) * rstd * gv.x + bv.x;
            float r1 = (xv.y - mean) * rstd * gv.y + bv.y;
            if (act != ACT_NONE) {
                r0 = apply_act(bf16_round(r0), act);
                r1 = apply_act(bf16_round(r1), act);
            }
            o[j] = pack_bf16x2(r0, r1);
        }
        yr[lane + 32 * k] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// HF LlamaRMSNorm (transformers 4.31 modeling_llama.py): variance in fp32, normalised value cast to bf16,
// then multiplied by the bf16 weight.
__global__ void rmsnorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const bf16* __restrict__ gamma,
                               long long rows, int D, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    const int nvec = D >> 3;
    float ss = 0.f;
    for (int i = lane; i < nvec; i += 32) {
        uint4 q = xr[i];
        float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    uint4* yr = reinterpret_cast<uint4*>(y + row * D);
    for (int i = lane; i < nvec; i += 32) {
        uint4 q = xr[i], g = g4[i];
        uint32_t xi[4] = {q.x, q.y, q.z, q.w}, gi[4] = {g.x, g.y, g.z, g.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]);
            o[j] = pack_bf16x2(gv.x * bf16_round(xv.x * rstd), gv.y * bf16_round(xv.y * rstd));
        }
        yr[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Few rows (decode steps: one row per sample): one CTA of 256 threads per row instead of one warp, so the row is
// covered by a single pass of 16-byte loads per thread and the kernel sits at the launch floor.
__global__ void __launch_bounds__(256) rmsnorm_row_cta_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                              const bf16* __restrict__ gamma, int D, float eps) {
    __shared__ float red[8];
    pdl_wait_then_launch();
    const long long row = blockIdx.x;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    const int nvec = D >> 3;
    constexpr int MAXV = 4;  // D <= 8192
    uint4 q[MAXV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int i = threadIdx.x + k * 256;
        q[k] = i < nvec ? xr[i] : make_uint4(0, 0, 0, 0);
        float2 a = unpack_bf16x2(q[k].x), b = unpack_bf16x2(q[k].y), c = unpack_bf16x2(q[k].z), d = unpack_bf16x2(q[k].w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / (float)D + eps);
    const uint4* g4 = reinterpret_cast<const uint4*>(gamma);
    uint4* yr = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i >= nvec) break;
        const uint4 g = g4[i];
        uint32_t xi[4] = {q[k].x, q[k].y, q[k].z, q[k].w}, gi[4] = {g.x, g.y, g.z, g.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 xv = unpack_bf16x2(xi[j]), gv = unpack_bf16x2(gi[j]);
            o[j] = pack_bf16x2(gv.x * bf16_round(xv.x * rstd), gv.y * bf16_round(xv.y * rstd));
        }
        yr[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void add_bcast_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                 long long nvec, long long pvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        uint4 x = a[i], y = b[pvec ? (i % pvec) : i];
        uint32_t xi[4] = {x.x, x.y, x.z, x.w}, yi[4] = {y.x, y.y, y.z, y.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 p = unpack_bf16x2(xi[j]), q = unpack_bf16x2(yi[j]);
            o[j] = pack_bf16x2(p.x + q.x, p.y + q.y);
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void silu_mul_kernel(const bf16* __restrict__ gu, bf16* __restrict__ out, long long rows, int F, int interleaved) {
    pdl_wait_then_launch();
    const int fvec = F >> 3;
    const long long total = rows * fvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / fvec;
        const int c = (int)(i % fvec);
        // interleaved: gate / up columns alternate in blocks of 8 (the row order of ivlm_decode_linear's SWIGLU epilogue)
        uint4 g = reinterpret_cast<const uint4*>(gu + r * 2 * F)[interleaved ? 2 * c : c];
        uint4 u = interleaved ? reinterpret_cast<const uint4*>(gu + r * 2 * F)[2 * c + 1]
                              : reinterpret_cast<const uint4*>(gu + r * 2 * F + F)[c];
        uint32_t gi[4] = {g.x, g.y, g.z, g.w}, ui[4] = {u.x, u.y, u.z, u.w}, o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 gv = unpack_bf16x2(gi[j]), uv = unpack_bf16x2(ui[j]);
            o[j] = pack_bf16x2(bf16_round(apply_act(gv.x, ACT_SILU)) * uv.x, bf16_round(apply_act(gv.y, ACT_SILU)) * uv.y);
        }
        reinterpret_cast<uint4*>(out + r * F)[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void fill_rows_kernel(bf16* __restrict__ out, long long ld, const int* __restrict__ rows, int n_rows,
                                 const bf16* __restrict__ vec, int nvec) {
    const long long total = (long long)n_rows * nvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / nvec), c = (int)(i % nvec);
        reinterpret_cast<uint4*>(out + (long long)rows[r] * ld)[c] = __ldg(reinterpret_cast<const uint4*>(vec) + c);
    }
}

__global__ void finalize_kernel(const float* __restrict__ acc, bf16* __restrict__ out, const bf16* __restrict__ bias,
                                const bf16* __restrict__ res, long long rows, int N, int act) {
    const long long total = rows * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % N);
        float x = acc[i];
        if (bias) x += __bfloat162float(bias[c]);
        x = bf16_round(x);
        if (act != ACT_NONE) x = bf16_round(apply_act(x, act));
        if (res) x += __bfloat162float(res[i]);
        out[i] = __float2bfloat16_rn(x);
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __bfloat162float(x[i]);
}

// ------------------------------------------------------------------------------------------------ im2col
// One thread per output element pair is overkill; one thread per (row, k) element with k fastest keeps the
// stores coalesced and the loads contiguous along dx.
__global__ void im2col_patch_kernel(const bf16* __restrict__ img, bf16* __restrict__ cols, int N, int C, int H, int W,
                                    int p, int ldk) {
    const int gw = W / p, gh = H / p;
    const long long total = (long long)N * gh * gw * ldk;
    const int kk = C * p * p;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % ldk);
        const long long row = i / ldk;
        bf16 v = __float2bfloat16_rn(0.f);
        if (k < kk) {
            const int dx = k % p, dy = (k / p) % p, c = k / (p * p);
            const int px = (int)(row % gw), py = (int)((row / gw) % gh);
            const long long n = row / ((long long)gw * gh);
            v = img[((n * C + c) * H + (py * p + dy)) * (long long)W + (px * p + dx)];
        }
        cols[i] = v;
    }
}

// Patch sizes that are multiples of 8 (SAM: 16): a (patch row, channel, dy) run of the operand is p contiguous bf16 in the image
// as well, so a thread moves 16 bytes; the grid's y / z carry (patch row, image) and the rest decodes with 32-bit arithmetic (the
// element-wise kernel above spends ~400 instructions per 2-byte element on 64-bit divisions: 0.22 ms per 8 views).
__global__ void __launch_bounds__(256)
im2col_patch_vec8_kernel(const bf16* __restrict__ img, bf16* __restrict__ cols, int C, int H, int W, int p, int ldk) {
    const int gw = W / p, gh = H / p, p8 = p >> 3;
    const int per_row = gw * C * p * p8;                      // 16-byte units of one row of patches
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_row) return;
    const int py = blockIdx.y;
    const long long n = blockIdx.z;
    int r = i;
    const int d8 = r % p8; r /= p8;
    const int dy = r % p; r /= p;
    const int c = r % C;
    const int px = r / C;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + ((n * C + c) * H + (py * p + dy)) * (long long)W + px * p + d8 * 8));
    const long long row = (n * gh + py) * gw + px;
    *reinterpret_cast<uint4*>(cols + row * ldk + (c * p + dy) * p + d8 * 8) = v;
}
__global__ void zero_tail_kernel(bf16* __restrict__ cols, long long rows, int kk, int ldk) {   // operand columns [kk, ldk) when ldk > kk
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tail = ldk - kk;
    if (i < rows * tail) cols[(i / tail) * ldk + kk + (int)(i % tail)] = __float2bfloat16_rn(0.f);
}

// 3x3, pad 1, token-major input [N,H,W,C]; 8 channels (16 B) per thread.
__global__ void im2col_3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ cols, int N, int H, int W, int C) {
    const int cvec = C >> 3;
    const long long total = (long long)N * H * W * 9 * cvec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cvec);
        const int tap = (int)((i / cvec) % 9);
        const long long row = i / (9LL * cvec);
        const int xx = (int)(row % W), yy = (int)((row / W) % H);
        const long long n = row / ((long long)W * H);
        const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (sy >= 0 && sy < H && sx >= 0 && sx < W)
            v = reinterpret_cast<const uint4*>(x + ((n * H + sy) * (long long)W + sx) * C)[cv];
        reinterpret_cast<uint4*>(cols + row * 9LL * C + (long long)tap * C)[cv] = v;
    }
}

// ------------------------------------------------------------------------------------------------ LLaVA glue
// out row (b, j): j < pos -> embed[ids[b,j]]; pos <= j < pos+n_img -> img[b, j-pos]; else embed[ids[b, j-n_img+1]].
// pos = index of the single IMAGE_TOKEN_INDEX (-200) in row b.
__global__ void embed_splice_kernel(const bf16* __restrict__ embed, const int* __restrict__ ids,
                                    const bf16* __restrict__ img, bf16* __restrict__ out, int B, int L, int n_img, int D,
                                    int vocab) {
    const int S = L - 1 + n_img;
    const long long row = blockIdx.x;
    const int b = (int)(row / S), j = (int)(row % S);
    __shared__ int s_pos;
    if (threadIdx.x == 0) {
        int pos = L;
        for (int t = 0; t < L; ++t)
            if (ids[b * L + t] < 0) { pos = t; break; }
        s_pos = pos;
    }
    __syncthreads();
    const int pos = s_pos;
    const uint4* src;
    if (j < pos) {
        int id = ids[b * L + j];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        src = reinterpret_cast<const uint4*>(embed + (long long)id * D);
    } else if (j < pos + n_img) {
        src = reinterpret_cast<const uint4*>(img + ((long long)b * n_img + (j - pos)) * D);
    } else {
        int id = ids[b * L + (j - n_img + 1)];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        src = reinterpret_cast<const uint4*>(embed + (long long)id * D);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + row * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src[i];
}

__global__ void gather_rows_kernel(const bf16* __restrict__ x, const int* __restrict__ idx, bf16* __restrict__ out, int n,
                                   int D, int max_row) {
    pdl_wait_then_launch();
    const int r = blockIdx.x;
    int s = idx[r];
    if (max_row > 0) s = s < 0 ? 0 : (s >= max_row ? max_row - 1 : s);
    const uint4* src = reinterpret_cast<const uint4*>(x + (long long)s * D);
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)r * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src[i];
}

// neq[i * K + k] = 1 when row i of x and row k of ref differ in any bit.  grid (segments, n * K): every CTA compares one
// segment of one pair with 16-byte loads and leaves at the first mismatch it (or another CTA of the pair) has seen.
__global__ void __launch_bounds__(256) rows_differ_kernel(const uint4* __restrict__ x, const uint4* __restrict__ ref, int K,
                                                          long long row_vecs, long long seg_vecs, int* __restrict__ neq) {
    const int pair = blockIdx.y;
    const int i = pair / K, k = pair % K;
    const uint4* a = x + (long long)i * row_vecs;
    const uint4* b = ref + (long long)k * row_vecs;
    const long long v0 = (long long)blockIdx.x * seg_vecs, v1 = min(v0 + seg_vecs, row_vecs);
    volatile int* flag = neq + pair;
    for (long long base = v0; base < v1; base += 256 * 4) {
        bool diff = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long v = base + u * 256 + threadIdx.x;
            if (v < v1) {
                const uint4 p = __ldg(a + v), q = __ldg(b + v);
                diff |= (p.x != q.x) | (p.y != q.y) | (p.z != q.z) | (p.w != q.w);
            }
        }
        if (__syncthreads_or(diff)) {
            if (threadIdx.x == 0) *flag = 1;
            return;
        }
        if (*flag) return;  // another segment of this pair already differs
    }
}

// torch.argmax semantics: first maximal index. One CTA per row.
__global__ void argmax_kernel(const float* __restrict__ logits, int* __restrict__ out, int vocab, long long ld) {
    pdl_wait_then_launch();
    const float* row = logits + (long long)blockIdx.x * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __shared__ float sv[32];
    __shared__ int si[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = best; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) out[blockIdx.x] = bi;
    }
}

// ------------------------------------------------------------------------------------------------ bilinear
// PyTorch upsample_bilinear2d, align_corners=False: src = max((dst+0.5)*scale-0.5, 0), scale = in/out.
__global__ void bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int sh, int sw, int ch,
                                int cw, int dh, int dw) {
    const float sy = (float)ch / (float)dh, sx = (float)cw / (float)dw;
    const long long total = (long long)N * dh * dw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % dw), y = (int)((i / dw) % dh);
        const long long n = i / ((long long)dw * dh);
        const BilinearTap t = bilinear_tap(y, x, sy, sx, ch, cw, sw);
        dst[i] = bilinear_eval(t, src + n * (long long)sh * sw);
    }
}

// Row-blocked form for dw % 4 == 0 and 16-byte aligned planes: one thread per four consecutive outputs of one row -- no 64-bit
// divisions (the generic kernel spends ~400 instructions per element on them: 0.56 TB/s), one 16-byte streaming store per thread,
// the source rows (1/16 of the output at x4) stay in L1/L2.  Same tap arithmetic -> bit-identical to bilinear_kernel.
__global__ void __launch_bounds__(256)
bilinear_rows4_kernel(const float* __restrict__ src, float* __restrict__ dst, int sh, int sw, int ch, int cw, int dh, int dw) {
    const float sy = (float)ch / (float)dh, sx = (float)cw / (float)dw;
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;     // group of four output columns
    if (x4 * 4 >= dw) return;
    const int y = blockIdx.y;
    const long long n = blockIdx.z;
    const float* plane = src + n * (long long)sh * sw;
    float4 o;
    float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const BilinearTap t = bilinear_tap(y, x4 * 4 + j, sy, sx, ch, cw, sw);
        ov[j] = bilinear_eval(t, plane);
    }
    __stcs(reinterpret_cast<float4*>(dst + (n * dh + y) * (long long)dw) + x4, o);
}

// ------------------------------------------------------------------------------------------------ camera gate
// One CTA of 256 threads per (sample, view). Every nn.Linear output is rounded to bf16 like the reference.
__global__ void cam_gate_kernel(const bf16* __restrict__ cam, const bf16* __restrict__ emb, const bf16* __restrict__ w1,
                                const bf16* __restrict__ b1, const bf16* __restrict__ w2, const bf16* __restrict__ b2,
                                const bf16* __restrict__ wv, const bf16* __restrict__ bv, bf16* __restrict__ out, int V) {
    const int b = blockIdx.x / V, v = blockIdx.x % V;
    __shared__ float h1[128], h2[128], c[5];
    const int t = threadIdx.x;
    if (t < 5) c[t] = __bfloat162float(cam[(b * V + v) * 5 + t]);
    __syncthreads();
    if (t < 128) {
        float s = 0.f;
        for (int k = 0; k < 5; ++k) s += c[k] * __bfloat162float(w1[t * 5 + k]);
        s = bf16_round(s + __bfloat162float(b1[t]));
        h1[t] = fmaxf(s, 0.f);
    }
    __syncthreads();
    if (t < 128) {
        float s = 0.f;
        for (int k = 0; k < 128; ++k) s += h1[k] * __bfloat162float(w2[t * 128 + k]);
        s = bf16_round(s + __bfloat162float(b2[t]));
        h2[t] = fmaxf(s, 0.f);
    }
    __syncthreads();
    {
        const bf16* w = wv + ((long long)v * 256 + t) * 128;
        float s = 0.f;
        for (int k = 0; k < 128; ++k) s += h2[k] * __bfloat162float(w[k]);
        s = bf16_round(s + __bfloat162float(bv[v * 256 + t]));
        const float g = bf16_round(1.f / (1.f + __expf(-s)));
        out[((long long)b * V + v) * 256 + t] = __float2bfloat16_rn(__bfloat162float(emb[b * 256 + t]) * g);
    }
}

// In-place sigmoid of mask logits wherever the ground-truth mask is not the ignore label (InteractVLM.py:452-456:
// heat-map view types feed sigmoid-ed maps to the point-cloud affordance lift).
__global__ void sigmoid_where_kernel(float* __restrict__ x, const float* __restrict__ gt, float ignore, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (gt == nullptr || gt[i] != ignore) x[i] = 1.f / (1.f + expf(-x[i]));
}

// uint8 HWC image -> normalised, zero-padded bf16 CHW: out[n,c,y,x] = bf16((img[n,y,x,c] * pre - mean[c]) / std[c]) for
// y < H, x < W, else 0.  SAM: pre = 1, mean/std in 0..255 units, S = 1024 (run_demo.py:65-79); CLIP: pre = 1/255
// (CLIPImageProcessor rescale) and S = H = W = 224.  One thread per 8 output pixels of a row (16-byte store).
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int N, int H, int W, int S,
                                     float pre, float m0, float m1, float m2, float s0, float s1, float s2) {
    const long long total = (long long)N * 3 * S * (S / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xv = (int)(i % (S / 8));
        long long r = i / (S / 8);
        const int y = (int)(r % S);
        r /= S;
        const int c = (int)(r % 3), n = (int)(r / 3);
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = xv * 8 + j;
            v[j] = (y < H && x < W) ? ((float)img[(((long long)n * H + y) * W + x) * 3 + c] * pre - mean) / sd : 0.f;
        }
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(out + (((long long)n * 3 + c) * S + y) * S + xv * 8) = q;
    }
}

// Pillow's separable antialiased resize on uint8 HWC images (libImaging/Resample.c, 8bpc): one pass along x or y.
// out[o] = clip8((2^21 + sum_j in[first[o] + j] * k[o][j]) >> 22); coefficient windows come from the host
// (interactvlm_b200/resample.py).  One thread per output pixel (3 channels), consecutive threads along x.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, int ksize, int N, int H, int W, int OH, int OW, int vertical) {
    const long long total = (long long)N * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const int oy = (int)((i / OW) % OH);
        const int n = (int)(i / ((long long)OW * OH));
        const int o = vertical ? oy : ox;
        const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
        const int* k = kk + (long long)o * ksize;
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
        for (int j = 0; j < cnt; ++j) {
            const int y = vertical ? first + j : oy, x = vertical ? ox : first + j;
            const uint8_t* px = src + (((long long)n * H + y) * W + x) * 3;
            const int c = k[j];
            a0 += px[0] * c; a1 += px[1] * c; a2 += px[2] * c;
        }
        uint8_t* q = dst + (((long long)n * OH + oy) * OW + ox) * 3;
        q[0] = (uint8_t)min(max(a0 >> 22, 0), 255);
        q[1] = (uint8_t)min(max(a1 >> 22, 0), 255);
        q[2] = (uint8_t)min(max(a2 >> 22, 0), 255);
    }
}

static inline int grid_for(long long work, int block, int sms) {
    long long g = (work + block - 1) / block;
    long long cap = (long long)sms * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ivlm

using namespace ivlm;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define DONE()                          \
    h->launches++;                      \
    IVLM_CHECK_CUDA(cudaGetLastError()); \
    return IVLM_OK

extern "C" int ivlm_layernorm_bf16(ivlm_handle h, const void* x, void* y, const void* gamma, const void* beta,
                                   int64_t out_rows, int32_t D, float eps, const int32_t* row_map, int32_t act,
                                   void* stream) {
    IVLM_REQUIRE(h && D % 8 == 0 && out_rows > 0, "layernorm: D=%d must be a multiple of 8, rows>0", D);
    const int wpb = 8;
    const unsigned grid = (unsigned)((out_rows + wpb - 1) / wpb);
#define IVLM_LN_REG(V)                                                                                                  \
    layernorm_reg_kernel<V><<<grid, wpb * 32, 0, STREAM>>>((const bf16*)x, (bf16*)y, (const bf16*)gamma, (const bf16*)beta, \
                                                           out_rows, eps, row_map, act)
    if (D == 1280) IVLM_LN_REG(5);
    else if (D == 1024) IVLM_LN_REG(4);
    else if (D == 256) IVLM_LN_REG(1);
    else
        layernorm_kernel<<<grid, wpb * 32, 0, STREAM>>>((const bf16*)x, (bf16*)y, (const bf16*)gamma, (const bf16*)beta,
                                                        out_rows, D, eps, row_map, act);
#undef IVLM_LN_REG
    DONE();
}
extern "C" int ivlm_rmsnorm_bf16(ivlm_handle h, const void* x, void* y, const void* gamma, int64_t rows, int32_t D,
                                 float eps, void* stream) {
    IVLM_REQUIRE(h && D % 8 == 0 && rows > 0, "rmsnorm: D=%d must be a multiple of 8, rows>0", D);
    // one CTA per row whenever the row fits its registers: a single pass of 16-byte loads per thread.  (The warp-per-row kernel walks
    // the row twice with one dependent load per iteration: 21 us per 2632 x 5120 prefill launch against 9 us of traffic.)
    if (D <= 8192 && rows <= 0x7fffffffLL) {
        IVLM_CHECK_CUDA(launch_k(h, rmsnorm_row_cta_kernel, dim3((unsigned)rows), dim3(256), 0, STREAM, (const bf16*)x, (bf16*)y,
                                 (const bf16*)gamma, D, eps));
    } else {
        const int wpb = 8;
        rmsnorm_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, STREAM>>>((const bf16*)x, (bf16*)y,
                                                                                    (const bf16*)gamma, rows, D, eps);
    }
    DONE();
}
extern "C" int ivlm_add_bcast_bf16(ivlm_handle h, const void* a, const void* b, void* out, int64_t n, int64_t period,
                                   void* stream) {
    IVLM_REQUIRE(h && n % 8 == 0 && period % 8 == 0, "add_bcast: n and period must be multiples of 8");
    add_bcast_kernel<<<grid_for(n / 8, 256, h->num_sms), 256, 0, STREAM>>>((const uint4*)a, (const uint4*)b, (uint4*)out,
                                                                           n / 8, period / 8);
    DONE();
}
extern "C" int ivlm_silu_mul_bf16(ivlm_handle h, const void* gate_up, void* out, int64_t rows, int32_t F, int32_t interleaved,
                                  void* stream) {
    IVLM_REQUIRE(h && F % 8 == 0, "silu_mul: F must be a multiple of 8");
    IVLM_CHECK_CUDA(launch_k(h, silu_mul_kernel, dim3(grid_for(rows * (F / 8), 256, h->num_sms)), dim3(256), 0, STREAM,
                             (const bf16*)gate_up, (bf16*)out, (long long)rows, F, (int)(interleaved ? 1 : 0)));
    DONE();
}
extern "C" int ivlm_fill_rows_bf16(ivlm_handle h, void* out, int64_t ld, const int32_t* rows, int32_t n_rows, const void* vec,
                                   int32_t N, void* stream) {
    IVLM_REQUIRE(h && out && rows && vec && n_rows > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "fill_rows: bad arguments");
    fill_rows_kernel<<<grid_for((long long)n_rows * (N / 8), 256, h->num_sms), 256, 0, STREAM>>>((bf16*)out, ld, rows, n_rows,
                                                                                               (const bf16*)vec, N / 8);
    DONE();
}
extern "C" int ivlm_finalize_f32_bf16(ivlm_handle h, const float* acc, void* out, const void* bias, const void* residual,
                                      int64_t rows, int32_t N, int32_t act, void* stream) {
    IVLM_REQUIRE(h && rows > 0 && N > 0, "finalize: empty");
    finalize_kernel<<<grid_for(rows * N, 256, h->num_sms), 256, 0, STREAM>>>(acc, (bf16*)out, (const bf16*)bias,
                                                                            (const bf16*)residual, rows, N, act);
    DONE();
}
extern "C" int ivlm_cast_f32_bf16(ivlm_handle h, const float* x, void* y, int64_t n, void* stream) {
    IVLM_REQUIRE(h && n > 0, "cast: empty");
    cast_f32_bf16_kernel<<<grid_for(n, 256, h->num_sms), 256, 0, STREAM>>>(x, (bf16*)y, n);
    DONE();
}
extern "C" int ivlm_cast_bf16_f32(ivlm_handle h, const void* x, float* y, int64_t n, void* stream) {
    IVLM_REQUIRE(h && n > 0, "cast: empty");
    cast_bf16_f32_kernel<<<grid_for(n, 256, h->num_sms), 256, 0, STREAM>>>((const bf16*)x, y, n);
    DONE();
}
extern "C" int ivlm_im2col_patch_bf16(ivlm_handle h, const void* img, void* cols, int32_t N, int32_t C, int32_t H,
                                      int32_t W, int32_t p, int32_t ldk, void* stream) {
    IVLM_REQUIRE(h && H % p == 0 && W % p == 0 && ldk >= C * p * p && ldk % 8 == 0, "im2col_patch: bad geometry");
    const long long total = (long long)N * (H / p) * (W / p) * ldk;
    if (p % 8 == 0 && W % 8 == 0 && H / p <= 65535 && N <= 65535 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(cols) & 15) == 0) {
        const int per_row = (W / p) * C * p * (p / 8);
        im2col_patch_vec8_kernel<<<dim3((per_row + 255) / 256, H / p, N), 256, 0, STREAM>>>((const bf16*)img, (bf16*)cols, C, H, W, p, ldk);
        if (ldk > C * p * p) {
            const long long rows = (long long)N * (H / p) * (W / p), n_tail = rows * (ldk - C * p * p);
            zero_tail_kernel<<<(unsigned)((n_tail + 255) / 256), 256, 0, STREAM>>>((bf16*)cols, rows, C * p * p, ldk);
        }
    } else {
        im2col_patch_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, STREAM>>>((const bf16*)img, (bf16*)cols, N, C, H, W, p,
                                                                                 ldk);
    }
    DONE();
}
extern "C" int ivlm_im2col_3x3_bf16(ivlm_handle h, const void* x, void* cols, int32_t N, int32_t H, int32_t W, int32_t C,
                                    void* stream) {
    IVLM_REQUIRE(h && C % 8 == 0, "im2col_3x3: C must be a multiple of 8");
    const long long total = (long long)N * H * W * 9 * (C / 8);
    im2col_3x3_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, STREAM>>>((const bf16*)x, (bf16*)cols, N, H, W, C);
    DONE();
}
extern "C" int ivlm_embed_splice_bf16(ivlm_handle h, const void* embed, const int32_t* ids, const void* img_feats,
                                      void* out, int32_t B, int32_t L, int32_t n_img, int32_t D, int32_t vocab,
                                      void* stream) {
    IVLM_REQUIRE(h && D % 8 == 0 && B > 0 && L > 0, "embed_splice: bad shape");
    embed_splice_kernel<<<B * (L - 1 + n_img), 128, 0, STREAM>>>((const bf16*)embed, ids, (const bf16*)img_feats,
                                                                  (bf16*)out, B, L, n_img, D, vocab);
    DONE();
}
extern "C" int ivlm_embed_gather_bf16(ivlm_handle h, const void* embed, const int32_t* ids, void* out, int32_t n,
                                      int32_t D, int32_t vocab, void* stream) {
    IVLM_REQUIRE(h && D % 8 == 0 && n > 0, "embed_gather: bad shape");
    IVLM_CHECK_CUDA(launch_k(h, gather_rows_kernel, dim3(n), dim3(128), 0, STREAM, (const bf16*)embed, (const int*)ids,
                             (bf16*)out, (int)n, (int)D, (int)vocab));
    DONE();
}
extern "C" int ivlm_gather_rows_bf16(ivlm_handle h, const void* x, const int32_t* idx, void* out, int32_t n, int32_t D,
                                     void* stream) {
    IVLM_REQUIRE(h && D % 8 == 0 && n > 0, "gather_rows: bad shape");
    IVLM_CHECK_CUDA(launch_k(h, gather_rows_kernel, dim3(n), dim3(128), 0, STREAM, (const bf16*)x, (const int*)idx,
                             (bf16*)out, (int)n, (int)D, 0));
    DONE();
}
extern "C" int ivlm_rows_differ(ivlm_handle h, const void* x, int32_t n, const void* ref, int32_t K, int64_t row_bytes,
                                int32_t* neq, void* stream) {
    IVLM_REQUIRE(h && x && ref && neq && n > 0 && K > 0 && row_bytes > 0 && row_bytes % 16 == 0,
                 "rows_differ: need n, K > 0 and a row size that is a multiple of 16 bytes");
    IVLM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ref) & 15) == 0,
                 "rows_differ: 16-byte aligned rows");
    const long long row_vecs = row_bytes / 16;
    int segs = (int)std::min<long long>(64, (row_vecs + 4095) / 4096);
    const long long seg_vecs = (row_vecs + segs - 1) / segs;
    IVLM_CHECK_CUDA(cudaMemsetAsync(neq, 0, sizeof(int) * (size_t)n * K, STREAM));
    rows_differ_kernel<<<dim3(segs, n * K), 256, 0, STREAM>>>((const uint4*)x, (const uint4*)ref, K, row_vecs, seg_vecs, neq);
    DONE();
}
extern "C" int ivlm_argmax_f32(ivlm_handle h, const float* logits, int32_t* out, int32_t B, int32_t vocab, int64_t ld,
                               void* stream) {
    IVLM_REQUIRE(h && B > 0 && vocab > 0, "argmax: empty");
    IVLM_CHECK_CUDA(launch_k(h, argmax_kernel, dim3(B), dim3(1024), 0, STREAM, logits, (int*)out, (int)vocab, (long long)ld));
    DONE();
}
extern "C" int ivlm_bilinear_f32(ivlm_handle h, const float* src, float* dst, int32_t N, int32_t sh, int32_t sw,
                                 int32_t crop_h, int32_t crop_w, int32_t dh, int32_t dw, void* stream) {
    IVLM_REQUIRE(h && crop_h <= sh && crop_w <= sw && crop_h > 0 && crop_w > 0 && N > 0, "bilinear: bad geometry");
    if (dw % 4 == 0 && dh <= 65535 && N <= 65535 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int tx = dw / 4 >= 256 ? 256 : ((dw / 4 + 31) / 32) * 32;
        bilinear_rows4_kernel<<<dim3((dw / 4 + tx - 1) / tx, dh, N), tx, 0, STREAM>>>(src, dst, sh, sw, crop_h, crop_w, dh, dw);
    } else {
        bilinear_kernel<<<grid_for((long long)N * dh * dw, 256, h->num_sms), 256, 0, STREAM>>>(src, dst, N, sh, sw, crop_h,
                                                                                             crop_w, dh, dw);
    }
    DONE();
}
extern "C" int ivlm_cam_gate_bf16(ivlm_handle h, const void* cam, const void* emb, const void* w1, const void* b1,
                                  const void* w2, const void* b2, const void* wv, const void* bv, void* out, int32_t B,
                                  int32_t V, void* stream) {
    IVLM_REQUIRE(h && B > 0 && V > 0, "cam_gate: empty");
    cam_gate_kernel<<<B * V, 256, 0, STREAM>>>((const bf16*)cam, (const bf16*)emb, (const bf16*)w1, (const bf16*)b1,
                                               (const bf16*)w2, (const bf16*)b2, (const bf16*)wv, (const bf16*)bv,
                                               (bf16*)out, V);
    DONE();
}

extern "C" int ivlm_sigmoid_where_f32(ivlm_handle h, float* x, const float* gt, float ignore_value, int64_t n, void* stream) {
    IVLM_REQUIRE(h && x && n > 0, "sigmoid_where: bad arguments");
    sigmoid_where_kernel<<<grid_for(n, 256, h->num_sms), 256, 0, STREAM>>>(x, gt, ignore_value, n);
    DONE();
}

extern "C" int ivlm_preprocess_u8_bf16(ivlm_handle h, const uint8_t* img, void* out, int32_t N, int32_t H, int32_t W, int32_t S,
                                       float pre_scale, const float* mean3_h, const float* std3_h, void* stream) {
    IVLM_REQUIRE(h && img && out && mean3_h && std3_h && N > 0 && H > 0 && W > 0 && H <= S && W <= S && S % 8 == 0,
                 "preprocess: bad arguments (N=%d H=%d W=%d S=%d)", N, H, W, S);
    const long long total = (long long)N * 3 * S * (S / 8);
    preprocess_u8_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, STREAM>>>(img, (bf16*)out, N, H, W, S, pre_scale, mean3_h[0],
                                                                             mean3_h[1], mean3_h[2], std3_h[0], std3_h[1], std3_h[2]);
    DONE();
}

extern "C" int ivlm_resample_u8(ivlm_handle h, const uint8_t* src, uint8_t* dst, const int32_t* bounds, const int32_t* coeffs,
                                int32_t ksize, int32_t N, int32_t H, int32_t W, int32_t OH, int32_t OW, int32_t vertical,
                                void* stream) {
    IVLM_REQUIRE(h && src && dst && bounds && coeffs && N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && ksize > 0,
                 "resample: bad arguments");
    IVLM_REQUIRE(vertical ? (OW == W) : (OH == H), "resample: a pass changes one dimension only");
    const long long total = (long long)N * OH * OW;
    resample_u8_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, STREAM>>>(src, dst, bounds, coeffs, ksize, N, H, W, OH, OW,
                                                                           vertical);
    DONE();
}
