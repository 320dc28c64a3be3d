// Render-Localise-Lift: multi-view 2D masks -> per-vertex contact probabilities.
//
// The reference scatters every valid pixel into its 3 vertices with atomics after copying 150 MB of maps to
// the GPU per sample (components.py:220-277).  Here the maps are inverted ONCE into a per-(view,vertex) CSR of
// (pixel, weight) entries, and the lift is a deterministic gather: no atomics, int32 indices, ~18 MB of map
// data shared by the whole batch.  Entries are ordered (corner k, pixel); a warp strides each (view, vertex) segment and
// folds its lanes' partial sums with shuffles (fp32; within 1e-6 of the reference's sequential scatter_add_ order).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "runtime.h"

struct ivlm_lift_map {
    int V = 0, H = 0, W = 0, n = 0;
    long long nnz = 0;
    int* row_ptr = nullptr;  // device [n*V + 1], vertex-major: row = vertex * V + view
    int* pix = nullptr;      // device [nnz] pixel index inside the view
    float* wgt = nullptr;    // device [nnz] barycentric weight (nullptr => unit weights)
};

struct ivlm_csr {
    int rows = 0, cols = 0;
    long long nnz = 0;
    int* row_ptr = nullptr;
    int* col = nullptr;
    float* val = nullptr;
};

namespace ivlm {

// One CTA (4 warps) per (vertex, chunk of BC samples).  The CSR is VERTEX-major (row = vertex * V + view), so the V segments of
// a vertex are adjacent; the CTA walks them view by view with ALL 128 threads striding a segment -- coalesced loads of pix /
// wgt, each entry loaded ONCE and applied to all samples of the chunk -- which keeps the four warps equally busy whatever the
// split of a vertex's pixels over the views (one warp per view left three warps idle behind the longest segment: 6890 blocks
// in 9 waves paced by their slowest warp).  Per-thread partial sums (registers, one set per view) are folded with shuffles and
// then across the warps in warp order: deterministic, the grouping depends only on the map.  Views are combined in view
// order like the reference's loop (components.py:235-262).  LOWRES: `src` holds the mask decoder's low-res logits
// [B,V,sh,sw] and the x(H/sh) bilinear of Sam.postprocess_masks (sam.py:161-165) is evaluated per entry with the arithmetic
// of bilinear_kernel (bit-identical to lifting the materialised 1024^2 logits, which are then never read: 1.05 MB instead of
// 16.8 MB per sample).
constexpr int LIFT_BC = 8, LIFT_WARPS = 4, LIFT_MAX_VIEWS = 8;

template <int MODE, bool LOWRES, int BC>
__global__ void __launch_bounds__(LIFT_WARPS * 32)
lift_warp_kernel(const int* __restrict__ row_ptr, const int* __restrict__ pix, const float* __restrict__ wgt,
                 const float* __restrict__ src, float* __restrict__ contact, int B, int V, int n, int H, int W, int sh, int sw,
                 float thr) {
    const int vtx = blockIdx.x, b0 = blockIdx.y * BC;
    const int nb = min(BC, B - b0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_votes[LIFT_MAX_VIEWS][LIFT_WARPS][BC], s_cnt[LIFT_MAX_VIEWS][LIFT_WARPS][BC];
    __shared__ int s_rp[LIFT_MAX_VIEWS + 1];
    if ((int)threadIdx.x <= V) s_rp[threadIdx.x] = row_ptr[(long long)vtx * V + threadIdx.x];
    __syncthreads();
    const long long plane = LOWRES ? (long long)sh * sw : (long long)H * W;
    const float sy = (float)sh / (float)H, sx = (float)sw / (float)W;
    for (int v = 0; v < V; ++v) {
        const int e0 = s_rp[v], e1 = s_rp[v + 1];
        float votes[BC], cnt[BC];
#pragma unroll
        for (int s = 0; s < BC; ++s) votes[s] = cnt[s] = 0.f;
        // sample slots past the batch end alias the last valid sample: every load below is unconditional, so all BC (x4 taps)
        // gathers of an entry are in flight together instead of one L2 round trip after the other
        const float* ms[BC];
#pragma unroll
        for (int s = 0; s < BC; ++s) ms[s] = src + ((long long)(b0 + min(s, nb - 1)) * V + v) * plane;
        for (int e = e0 + (int)threadIdx.x; e < e1; e += LIFT_WARPS * 32) {
            const int p = __ldg(pix + e);
            const float w = (MODE == IVLM_LIFT_POINTS) ? 1.f : __ldg(wgt + e);
            float x[BC];
            if (LOWRES) {
                const BilinearTap t = bilinear_tap(p / W, p % W, sy, sx, sh, sw, sw);
                float a[BC], b[BC], c[BC], d[BC];
#pragma unroll
                for (int s = 0; s < BC; ++s) {
                    a[s] = __ldg(ms[s] + t.o00); b[s] = __ldg(ms[s] + t.o01);
                    c[s] = __ldg(ms[s] + t.o10); d[s] = __ldg(ms[s] + t.o11);
                }
#pragma unroll
                for (int s = 0; s < BC; ++s) x[s] = bilinear_combine(t, a[s], b[s], c[s], d[s]);
            } else {
#pragma unroll
                for (int s = 0; s < BC; ++s) x[s] = __ldg(ms[s] + p);
            }
            if (MODE != IVLM_LIFT_OBJECT_MESH) cnt[0] += w;   // the weight sum does not depend on the sample
#pragma unroll
            for (int s = 0; s < BC; ++s) {
                if (MODE == IVLM_LIFT_POINTS) {
                    votes[s] += x[s];
                } else {
                    const float xc = (MODE == IVLM_LIFT_HUMAN) ? fminf(fmaxf(x[s], -20.f), 20.f) : x[s];
                    const float pr = 1.f / (1.f + expf(-xc));
                    if (MODE == IVLM_LIFT_OBJECT_MESH) {
                        const bool on = pr > thr;
                        votes[s] += on ? w * pr : 0.f;
                        cnt[s] += on ? w : 0.f;
                    } else {
                        votes[s] += w * pr;
                    }
                }
            }
        }
        if (e1 - e0 > 0) {   // CTA-uniform: empty segments skip the shuffles
            if (MODE != IVLM_LIFT_OBJECT_MESH) cnt[0] = warp_sum(cnt[0]);
#pragma unroll
            for (int s = 0; s < BC; ++s) {
                votes[s] = warp_sum(votes[s]);
                if (MODE == IVLM_LIFT_OBJECT_MESH) cnt[s] = warp_sum(cnt[s]);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < BC; ++s) {
                s_votes[v][warp][s] = votes[s];
                s_cnt[v][warp][s] = (MODE == IVLM_LIFT_OBJECT_MESH) ? cnt[s] : cnt[0];
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
        const int s = threadIdx.x;
        float pred = 0.f, nviews = 0.f;
        for (int v = 0; v < V; ++v) {
            float c = 0.f, vt = 0.f;
#pragma unroll
            for (int w = 0; w < LIFT_WARPS; ++w) { c += s_cnt[v][w][s]; vt += s_votes[v][w][s]; }
            // votes are normalised only where the weight sum is positive; the rest is added as is (components.py:257-262)
            pred += (c > 0.f) ? vt / c : vt;
            nviews += (c > 0.f) ? 1.f : 0.f;
        }
        if (nviews > 0.f) pred = pred / nviews;
        if (MODE == IVLM_LIFT_HUMAN) pred = fminf(fmaxf(pred, 0.f), 1.f);
        contact[(long long)(b0 + s) * n + vtx] = pred;
    }
}

template <bool LOWRES, int BC>
static cudaError_t launch_lift_bc(const ivlm_lift_map* m, const float* src, float* contact, int B, int mode, float thr, int sh,
                                  int sw, cudaStream_t st) {
    const dim3 grid((unsigned)m->n, (unsigned)((B + BC - 1) / BC)), block(LIFT_WARPS * 32);
#define IVLM_LIFT_LAUNCH(MODE)                                                                                                  \
    lift_warp_kernel<MODE, LOWRES, BC><<<grid, block, 0, st>>>(m->row_ptr, m->pix, m->wgt, src, contact, B, m->V, m->n, m->H, \
                                                                m->W, sh, sw, thr)
    if (mode == IVLM_LIFT_HUMAN) IVLM_LIFT_LAUNCH(IVLM_LIFT_HUMAN);
    else if (mode == IVLM_LIFT_OBJECT_MESH) IVLM_LIFT_LAUNCH(IVLM_LIFT_OBJECT_MESH);
    else IVLM_LIFT_LAUNCH(IVLM_LIFT_POINTS);
#undef IVLM_LIFT_LAUNCH
    return cudaGetLastError();
}

// samples per CTA pass: the smallest of 1 / 2 / 4 / 8 that covers the batch (a map entry is loaded once per pass)
template <bool LOWRES>
static cudaError_t launch_lift(const ivlm_lift_map* m, const float* src, float* contact, int B, int mode, float thr, int sh,
                               int sw, cudaStream_t st) {
    if (B <= 1) return launch_lift_bc<LOWRES, 1>(m, src, contact, B, mode, thr, sh, sw, st);
    if (B <= 2) return launch_lift_bc<LOWRES, 2>(m, src, contact, B, mode, thr, sh, sw, st);
    if (B <= 4) return launch_lift_bc<LOWRES, 4>(m, src, contact, B, mode, thr, sh, sw, st);
    return launch_lift_bc<LOWRES, LIFT_BC>(m, src, contact, B, mode, thr, sh, sw, st);
}

__global__ void csr_spmv_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col,
                                const float* __restrict__ val, const float* __restrict__ x, float* __restrict__ y,
                                int B, int rows, int cols) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * rows) return;
    const int b = (int)(idx / rows), r = (int)(idx % rows);
    float acc = 0.f;
    for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) acc += val[e] * x[(long long)b * cols + col[e]];
    y[idx] = acc;
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& host) {
    const size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
    IVLM_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), bytes));
    if (!host.empty()) IVLM_CHECK_CUDA(cudaMemcpy(*dptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return IVLM_OK;
}

// In-place inclusive prefix sum of per-row counts held in row_ptr[1..]; refuses maps whose entry count leaves int32.
static int prefix_sum_checked(std::vector<int>& row_ptr, const char* who) {
    long long run = 0;
    for (size_t i = 1; i < row_ptr.size(); ++i) {
        run += row_ptr[i];
        IVLM_REQUIRE(run < (1LL << 31), "%s: %lld map entries exceed the int32 CSR range", who, run);
        row_ptr[i] = (int)run;
    }
    return IVLM_OK;
}

}  // namespace ivlm

using namespace ivlm;
extern "C" int ivlm_lift_free(ivlm_lift_map* m);
extern "C" int ivlm_csr_free(ivlm_csr* m);

extern "C" int ivlm_lift_build_mesh(ivlm_handle h, const int64_t* p2v, const float* bary, int32_t V, int32_t H, int32_t W,
                                    int32_t n, ivlm_lift_map** out) {
    IVLM_REQUIRE(h && p2v && bary && out && V > 0 && H > 0 && W > 0 && n > 0, "lift_build_mesh: bad arguments");
    IVLM_REQUIRE((long long)H * W < (1LL << 31), "lift_build_mesh: view too large for int32 pixel ids");
    const long long hw = (long long)H * W;
    std::vector<int> row_ptr((size_t)V * n + 1, 0);
    // pass 1: counts (a pixel votes only if all three ids are valid, components.py:241-245)
    // rows are vertex-major: row = vertex * V + view (the V segments of a vertex are adjacent)
    for (int v = 0; v < V; ++v) {
        const int64_t* pv = p2v + (long long)v * hw * 3;
        int* cnt = row_ptr.data() + 1 + v;
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = pv[i * 3], b = pv[i * 3 + 1], c = pv[i * 3 + 2];
            if (a < 0 || a >= n || b < 0 || b >= n || c < 0 || c >= n) continue;
            cnt[a * V]++; cnt[b * V]++; cnt[c * V]++;
        }
    }
    IVLM_TRY(prefix_sum_checked(row_ptr, "lift_build_mesh"));
    const long long nnz = row_ptr.back();
    std::vector<int> pix((size_t)nnz);
    std::vector<float> wgt((size_t)nnz);
    std::vector<int> cur(row_ptr.begin(), row_ptr.end() - 1);
    // pass 2: fill, corner-major then pixel order (= order of the reference's three scatter_add_ passes)
    for (int v = 0; v < V; ++v) {
        const int64_t* pv = p2v + (long long)v * hw * 3;
        const float* bw = bary + (long long)v * hw * 3;
        int* c = cur.data() + v;
        for (int k = 0; k < 3; ++k) {
            for (long long i = 0; i < hw; ++i) {
                const int64_t a = pv[i * 3], b = pv[i * 3 + 1], d = pv[i * 3 + 2];
                if (a < 0 || a >= n || b < 0 || b >= n || d < 0 || d >= n) continue;
                const int64_t vid = pv[i * 3 + k];
                const int e = c[vid * V]++;
                pix[e] = (int)i;
                wgt[e] = bw[i * 3 + k];
            }
        }
    }
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    ivlm_lift_map* m = new ivlm_lift_map();
    m->V = V; m->H = H; m->W = W; m->n = n; m->nnz = nnz;
    if (upload(&m->row_ptr, row_ptr) != IVLM_OK || upload(&m->pix, pix) != IVLM_OK || upload(&m->wgt, wgt) != IVLM_OK) {
        ivlm_lift_free(m);
        return IVLM_ERR_CUDA;
    }
    *out = m;
    return IVLM_OK;
}

extern "C" int ivlm_lift_build_points(ivlm_handle h, const int64_t* p2p, int32_t V, int32_t H, int32_t W, int32_t n,
                                      ivlm_lift_map** out) {
    IVLM_REQUIRE(h && p2p && out && V > 0 && H > 0 && W > 0 && n > 0, "lift_build_points: bad arguments");
    IVLM_REQUIRE((long long)H * W < (1LL << 31), "lift_build_points: view too large for int32 pixel ids");
    const long long hw = (long long)H * W;
    std::vector<int> row_ptr((size_t)V * n + 1, 0);
    for (int v = 0; v < V; ++v)
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = p2p[(long long)v * hw + i];
            if (a == -1) continue;  // components.py:327 `pixel_to_point_map != -1`
            IVLM_REQUIRE(a >= 0 && a < n, "lift_build_points: point id %lld out of range [0,%d)", (long long)a, n);
            row_ptr[(size_t)a * V + v + 1]++;
        }
    IVLM_TRY(prefix_sum_checked(row_ptr, "lift_build_points"));
    const long long nnz = row_ptr.back();
    std::vector<int> pix((size_t)nnz);
    std::vector<int> cur(row_ptr.begin(), row_ptr.end() - 1);
    for (int v = 0; v < V; ++v)
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = p2p[(long long)v * hw + i];
            if (a == -1) continue;
            pix[cur[(size_t)a * V + v]++] = (int)i;
        }
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    ivlm_lift_map* m = new ivlm_lift_map();
    m->V = V; m->H = H; m->W = W; m->n = n; m->nnz = nnz;
    if (upload(&m->row_ptr, row_ptr) != IVLM_OK || upload(&m->pix, pix) != IVLM_OK) {
        ivlm_lift_free(m);
        return IVLM_ERR_CUDA;
    }
    *out = m;
    return IVLM_OK;
}

extern "C" int ivlm_lift_free(ivlm_lift_map* m) {
    if (!m) return IVLM_OK;
    cudaFree(m->row_ptr);
    cudaFree(m->pix);
    cudaFree(m->wgt);
    delete m;
    return IVLM_OK;
}
extern "C" int64_t ivlm_lift_nnz(const ivlm_lift_map* m) { return m ? m->nnz : 0; }

extern "C" int ivlm_lift(ivlm_handle h, const ivlm_lift_map* m, const float* masks, float* contact, int32_t B,
                         int32_t mode, float thr, void* stream) {
    IVLM_REQUIRE(h && m && masks && contact && B > 0, "lift: bad arguments");
    IVLM_REQUIRE(mode == IVLM_LIFT_HUMAN || mode == IVLM_LIFT_OBJECT_MESH || mode == IVLM_LIFT_POINTS, "lift: unknown mode %d", mode);
    IVLM_REQUIRE(mode == IVLM_LIFT_POINTS || m->wgt != nullptr, "lift: mesh modes need a map built by ivlm_lift_build_mesh");
    IVLM_REQUIRE(m->V <= LIFT_MAX_VIEWS, "lift: at most %d views", LIFT_MAX_VIEWS);
    IVLM_CHECK_CUDA(launch_lift<false>(m, masks, contact, B, mode, thr, m->H, m->W, reinterpret_cast<cudaStream_t>(stream)));
    h->launches++;
    return IVLM_OK;
}

extern "C" int ivlm_lift_lowres(ivlm_handle h, const ivlm_lift_map* m, const float* lowres, int32_t sh, int32_t sw,
                                float* contact, int32_t B, int32_t mode, float thr, void* stream) {
    IVLM_REQUIRE(h && m && lowres && contact && B > 0 && sh > 0 && sw > 0, "lift_lowres: bad arguments");
    IVLM_REQUIRE(mode == IVLM_LIFT_HUMAN || mode == IVLM_LIFT_OBJECT_MESH || mode == IVLM_LIFT_POINTS, "lift_lowres: unknown mode %d", mode);
    IVLM_REQUIRE(mode == IVLM_LIFT_POINTS || m->wgt != nullptr, "lift_lowres: mesh modes need a map built by ivlm_lift_build_mesh");
    IVLM_REQUIRE(m->V <= LIFT_MAX_VIEWS, "lift_lowres: at most %d views", LIFT_MAX_VIEWS);
    IVLM_REQUIRE((long long)sh * sw < (1LL << 31), "lift_lowres: source plane too large");
    IVLM_CHECK_CUDA(launch_lift<true>(m, lowres, contact, B, mode, thr, sh, sw, reinterpret_cast<cudaStream_t>(stream)));
    h->launches++;
    return IVLM_OK;
}

extern "C" int ivlm_csr_build_dense(ivlm_handle h, const float* dense, int32_t rows, int32_t cols, ivlm_csr** out) {
    IVLM_REQUIRE(h && dense && out && rows > 0 && cols > 0, "csr_build_dense: bad arguments");
    std::vector<int> row_ptr((size_t)rows + 1, 0), col;
    std::vector<float> val;
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            const float x = dense[(size_t)r * cols + c];
            if (x != 0.f) { col.push_back(c); val.push_back(x); }
        }
        row_ptr[r + 1] = (int)col.size();
    }
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    ivlm_csr* m = new ivlm_csr();
    m->rows = rows; m->cols = cols; m->nnz = (long long)col.size();
    if (upload(&m->row_ptr, row_ptr) != IVLM_OK || upload(&m->col, col) != IVLM_OK || upload(&m->val, val) != IVLM_OK) {
        ivlm_csr_free(m);
        return IVLM_ERR_CUDA;
    }
    *out = m;
    return IVLM_OK;
}
extern "C" int ivlm_csr_free(ivlm_csr* m) {
    if (!m) return IVLM_OK;
    cudaFree(m->row_ptr);
    cudaFree(m->col);
    cudaFree(m->val);
    delete m;
    return IVLM_OK;
}
extern "C" int ivlm_csr_spmv(ivlm_handle h, const ivlm_csr* m, const float* x, float* y, int32_t B, void* stream) {
    IVLM_REQUIRE(h && m && x && y && B > 0, "csr_spmv: bad arguments");
    const long long total = (long long)B * m->rows;
    csr_spmv_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        m->row_ptr, m->col, m->val, x, y, B, m->rows, m->cols);
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
    return IVLM_OK;
}
