// Contact term of the pose refinement (optim/optimizer.py:80-96, ObjPose_Opt.contact_loss): the reference materialises
// the [N_obj, N_human] distance matrix with torch.cdist and the outer product of the two contact-probability vectors
// (2 x 0.8 GB for a 20 k-vertex object against SMPL-X) and autograd keeps both for the backward pass.  Here value and
// gradient come from one pass over the pairs, nothing of size N_obj x N_human is ever stored:
//   loss = sum_ij p_i q_j |o_i - h_j| / (sum_i p_i * sum_j q_j),   d loss / d o_i = p_i sum_j q_j (o_i - h_j) / |o_i - h_j| / (P Q)
// Grid: (object-vertex tiles, human-vertex splits); human vertices are staged through shared memory; each split writes
// its partial row sums to the workspace and a single finalising CTA adds them in a fixed order (deterministic).
#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int CL_THREADS = 128;   // object vertices per CTA
constexpr int CL_CHUNK = 512;     // human vertices per shared-memory chunk

__global__ void __launch_bounds__(CL_THREADS)
contact_pairs_kernel(const float* __restrict__ obj, const float* __restrict__ hum, const float* __restrict__ hum_prob,
                     int n_obj, int n_hum, int per_split, float4* __restrict__ partial /*[splits, n_obj]*/) {
    __shared__ float4 sh[CL_CHUNK];
    const int i = blockIdx.x * CL_THREADS + threadIdx.x;
    const int j0 = blockIdx.y * per_split, j1 = min(j0 + per_split, n_hum);
    const bool live = i < n_obj;
    const float ox = live ? obj[3 * i] : 0.f, oy = live ? obj[3 * i + 1] : 0.f, oz = live ? obj[3 * i + 2] : 0.f;
    float s = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    for (int base = j0; base < j1; base += CL_CHUNK) {
        const int n = min(CL_CHUNK, j1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += CL_THREADS) {
            const int j = base + k;
            sh[k] = make_float4(hum[3 * j], hum[3 * j + 1], hum[3 * j + 2], hum_prob[j]);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const float4 h = sh[k];
            const float dx = ox - h.x, dy = oy - h.y, dz = oz - h.z;
            const float d2 = dx * dx + dy * dy + dz * dz;
            const float inv = d2 > 0.f ? rsqrtf(d2) : 0.f;  // cdist's backward is 0 at zero distance
            s += h.w * (d2 * inv);
            const float w = h.w * inv;
            gx += w * dx; gy += w * dy; gz += w * dz;
        }
    }
    if (live) partial[(size_t)blockIdx.y * n_obj + i] = make_float4(s, gx, gy, gz);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(1024)
contact_finalize_kernel(const float4* __restrict__ partial, const float* __restrict__ obj_prob,
                        const float* __restrict__ hum_prob, int n_obj, int n_hum, int splits, float* __restrict__ loss,
                        float* __restrict__ grad) {
    __shared__ double red[32];
    double p = 0.0, q = 0.0, num = 0.0;
    for (int j = threadIdx.x; j < n_hum; j += blockDim.x) q += hum_prob[j];
    for (int i = threadIdx.x; i < n_obj; i += blockDim.x) {
        const float pi = obj_prob[i];
        p += pi;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += partial[(size_t)k * n_obj + i].x;
        num += (double)pi * s;
    }
    const double P = block_sum(p, red), Q = block_sum(q, red), N = block_sum(num, red);
    const double denom = P * Q;
    if (threadIdx.x == 0) loss[0] = (float)(N / denom);
    if (grad == nullptr) return;
    for (int i = threadIdx.x; i < n_obj; i += blockDim.x) {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int k = 0; k < splits; ++k) {
            const float4 t = partial[(size_t)k * n_obj + i];
            gx += t.y; gy += t.z; gz += t.w;
        }
        const float sc = (float)((double)obj_prob[i] / denom);
        grad[3 * i] = sc * gx; grad[3 * i + 1] = sc * gy; grad[3 * i + 2] = sc * gz;
    }
}

// Nearest neighbour (K = 1) of every query row in a target set, D <= 8 coordinates per point (ICP over points ++ normals,
// optim/icp/icp.py:187-196 -> pytorch3d knn_points).  One thread per query, targets staged through shared memory, squared
// Euclidean distance summed coordinate by coordinate; ties keep the lowest target index.
constexpr int KNN_THREADS = 128, KNN_CHUNK = 512, KNN_MAX_D = 8;

template <int D>
__global__ void __launch_bounds__(KNN_THREADS)
knn1_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, int m, int* __restrict__ idx,
            float* __restrict__ dist2) {
    __shared__ float sh[KNN_CHUNK * D];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    float q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = i < n ? x[(size_t)i * D + d] : 0.f;
    float best = INFINITY;
    int best_j = 0;
    for (int base = 0; base < m; base += KNN_CHUNK) {
        const int c = min(KNN_CHUNK, m - base);
        __syncthreads();
        for (int k = threadIdx.x; k < c * D; k += KNN_THREADS) sh[k] = y[(size_t)base * D + k];
        __syncthreads();
        for (int k = 0; k < c; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float t = q[d] - sh[k * D + d];
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            if (acc < best) { best = acc; best_j = base + k; }
        }
    }
    if (i < n) {
        idx[i] = best_j;
        if (dist2) dist2[i] = best;
    }
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_knn1(ivlm_handle h, const float* x, const float* y, int32_t n, int32_t m, int32_t D, int32_t* idx,
                         float* dist2, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && x && y && idx, "knn1: null argument");
    IVLM_REQUIRE(n > 0 && m > 0 && D >= 1 && D <= KNN_MAX_D, "knn1: need n, m > 0 and 1 <= D <= %d (got %d, %d, %d)", KNN_MAX_D, n, m, D);
    const dim3 grid((n + KNN_THREADS - 1) / KNN_THREADS);
#define IVLM_KNN(D_) case D_: knn1_kernel<D_><<<grid, KNN_THREADS, 0, stream>>>(x, y, n, m, idx, dist2); break;
    switch (D) {
        IVLM_KNN(1) IVLM_KNN(2) IVLM_KNN(3) IVLM_KNN(4) IVLM_KNN(5) IVLM_KNN(6) IVLM_KNN(7) IVLM_KNN(8)
    }
#undef IVLM_KNN
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 1;
    return IVLM_OK;
}

extern "C" int ivlm_contact_loss(ivlm_handle h, const float* obj_verts, const float* obj_prob, const float* hum_verts,
                                 const float* hum_prob, int32_t n_obj, int32_t n_hum, float* loss, float* grad_obj,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && obj_verts && obj_prob && hum_verts && hum_prob && loss, "contact_loss: null argument");
    IVLM_REQUIRE(n_obj > 0 && n_hum > 0, "contact_loss: empty vertex set (%d object, %d human)", n_obj, n_hum);
    IVLM_REQUIRE(h->ws != nullptr, "contact_loss: needs the handle's workspace (ivlm_set_workspace)");
    const int tiles = (n_obj + CL_THREADS - 1) / CL_THREADS;
    // enough CTAs for ~2 waves; never more splits than chunks, and the partials must fit the workspace
    int splits = (2 * h->num_sms + tiles - 1) / tiles;
    const int max_splits = (n_hum + CL_CHUNK - 1) / CL_CHUNK;
    splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
    const size_t avail = h->ws_bytes - IVLM_WS_COUNTER_BYTES;
    while (splits > 1 && (size_t)splits * n_obj * sizeof(float4) > avail) --splits;
    IVLM_REQUIRE((size_t)splits * n_obj * sizeof(float4) <= avail, "contact_loss: workspace too small for %d object vertices", n_obj);
    int per_split = (n_hum + splits - 1) / splits;
    per_split = (per_split + CL_CHUNK - 1) / CL_CHUNK * CL_CHUNK;
    splits = (n_hum + per_split - 1) / per_split;
    float4* partial = reinterpret_cast<float4*>(h->ws + IVLM_WS_COUNTER_BYTES);
    contact_pairs_kernel<<<dim3(tiles, splits), CL_THREADS, 0, stream>>>(obj_verts, hum_verts, hum_prob, n_obj, n_hum, per_split,
                                                                         partial);
    contact_finalize_kernel<<<1, 1024, 0, stream>>>(partial, obj_prob, hum_prob, n_obj, n_hum, splits, loss, grad_obj);
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 2;
    return IVLM_OK;
}
