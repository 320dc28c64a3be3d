// Render-Localise-Lift: multi-view 2D masks -> per-vertex contact probabilities.
//
// The reference scatters every valid pixel into its 3 vertices with atomics after copying 150 MB of maps to
// the GPU per sample (components.py:220-277).  Here the maps are inverted ONCE into a per-(view,vertex) CSR of
// (pixel, weight) entries, and the lift is a deterministic gather: no atomics, int32 indices, ~18 MB of map
// data shared by the whole batch.  Entries are ordered (corner k, pixel) and summed sequentially with separate
// multiply and add roundings, which is the order the reference's three scatter_add_ passes use on CPU.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "runtime.h"

struct ivlm_lift_map {
    int V = 0, H = 0, W = 0, n = 0;
    long long nnz = 0;
    int* row_ptr = nullptr;  // device [V*n + 1]
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

__global__ void lift_kernel(const int* __restrict__ row_ptr, const int* __restrict__ pix, const float* __restrict__ wgt,
                            const float* __restrict__ masks, float* __restrict__ contact, int B, int V, int n,
                            long long hw, int mode, float thr) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * n) return;
    const int b = (int)(idx / n), vtx = (int)(idx % n);
    float pred = 0.f, nviews = 0.f;
    for (int v = 0; v < V; ++v) {
        const float* m = masks + ((long long)b * V + v) * hw;
        const int e0 = row_ptr[v * n + vtx], e1 = row_ptr[v * n + vtx + 1];
        float votes = 0.f, cnt = 0.f;
        for (int e = e0; e < e1; ++e) {
            float x = __ldg(m + pix[e]);
            if (mode == IVLM_LIFT_POINTS) {
                votes = __fadd_rn(votes, x);
                cnt = __fadd_rn(cnt, 1.f);
                continue;
            }
            if (mode == IVLM_LIFT_HUMAN) x = fminf(fmaxf(x, -20.f), 20.f);
            const float p = 1.f / (1.f + expf(-x));
            if (mode == IVLM_LIFT_OBJECT_MESH && !(p > thr)) continue;
            const float w = wgt[e];
            votes = __fadd_rn(votes, __fmul_rn(w, p));
            cnt = __fadd_rn(cnt, w);
        }
        if (cnt > 0.f) {
            pred = __fadd_rn(pred, votes / cnt);
            nviews += 1.f;
        }
    }
    if (nviews > 0.f) pred = pred / nviews;
    if (mode == IVLM_LIFT_HUMAN) pred = fminf(fmaxf(pred, 0.f), 1.f);
    contact[idx] = pred;
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

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_lift_build_mesh(ivlm_handle h, const int64_t* p2v, const float* bary, int32_t V, int32_t H, int32_t W,
                                    int32_t n, ivlm_lift_map** out) {
    IVLM_REQUIRE(h && p2v && bary && out && V > 0 && H > 0 && W > 0 && n > 0, "lift_build_mesh: bad arguments");
    IVLM_REQUIRE((long long)H * W < (1LL << 31), "lift_build_mesh: view too large for int32 pixel ids");
    const long long hw = (long long)H * W;
    std::vector<int> row_ptr((size_t)V * n + 1, 0);
    // pass 1: counts (a pixel votes only if all three ids are valid, components.py:241-245)
    for (int v = 0; v < V; ++v) {
        const int64_t* pv = p2v + (long long)v * hw * 3;
        int* cnt = row_ptr.data() + (size_t)v * n + 1;
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = pv[i * 3], b = pv[i * 3 + 1], c = pv[i * 3 + 2];
            if (a < 0 || a >= n || b < 0 || b >= n || c < 0 || c >= n) continue;
            cnt[a]++; cnt[b]++; cnt[c]++;
        }
    }
    for (size_t i = 1; i < row_ptr.size(); ++i) row_ptr[i] += row_ptr[i - 1];
    const long long nnz = row_ptr.back();
    std::vector<int> pix((size_t)nnz);
    std::vector<float> wgt((size_t)nnz);
    std::vector<int> cur(row_ptr.begin(), row_ptr.end() - 1);
    // pass 2: fill, corner-major then pixel order (= order of the reference's three scatter_add_ passes)
    for (int v = 0; v < V; ++v) {
        const int64_t* pv = p2v + (long long)v * hw * 3;
        const float* bw = bary + (long long)v * hw * 3;
        int* c = cur.data() + (size_t)v * n;
        for (int k = 0; k < 3; ++k) {
            for (long long i = 0; i < hw; ++i) {
                const int64_t a = pv[i * 3], b = pv[i * 3 + 1], d = pv[i * 3 + 2];
                if (a < 0 || a >= n || b < 0 || b >= n || d < 0 || d >= n) continue;
                const int64_t vid = pv[i * 3 + k];
                const int e = c[vid]++;
                pix[e] = (int)i;
                wgt[e] = bw[i * 3 + k];
            }
        }
    }
    ivlm_lift_map* m = new ivlm_lift_map();
    m->V = V; m->H = H; m->W = W; m->n = n; m->nnz = nnz;
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    IVLM_TRY(upload(&m->row_ptr, row_ptr));
    IVLM_TRY(upload(&m->pix, pix));
    IVLM_TRY(upload(&m->wgt, wgt));
    *out = m;
    return IVLM_OK;
}

extern "C" int ivlm_lift_build_points(ivlm_handle h, const int64_t* p2p, int32_t V, int32_t H, int32_t W, int32_t n,
                                      ivlm_lift_map** out) {
    IVLM_REQUIRE(h && p2p && out && V > 0 && H > 0 && W > 0 && n > 0, "lift_build_points: bad arguments");
    const long long hw = (long long)H * W;
    std::vector<int> row_ptr((size_t)V * n + 1, 0);
    for (int v = 0; v < V; ++v)
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = p2p[(long long)v * hw + i];
            if (a == -1) continue;  // components.py:327 `pixel_to_point_map != -1`
            IVLM_REQUIRE(a >= 0 && a < n, "lift_build_points: point id %lld out of range [0,%d)", (long long)a, n);
            row_ptr[(size_t)v * n + 1 + a]++;
        }
    for (size_t i = 1; i < row_ptr.size(); ++i) row_ptr[i] += row_ptr[i - 1];
    const long long nnz = row_ptr.back();
    std::vector<int> pix((size_t)nnz);
    std::vector<int> cur(row_ptr.begin(), row_ptr.end() - 1);
    for (int v = 0; v < V; ++v)
        for (long long i = 0; i < hw; ++i) {
            const int64_t a = p2p[(long long)v * hw + i];
            if (a == -1) continue;
            pix[cur[(size_t)v * n + a]++] = (int)i;
        }
    ivlm_lift_map* m = new ivlm_lift_map();
    m->V = V; m->H = H; m->W = W; m->n = n; m->nnz = nnz;
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    IVLM_TRY(upload(&m->row_ptr, row_ptr));
    IVLM_TRY(upload(&m->pix, pix));
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
    IVLM_REQUIRE(mode == IVLM_LIFT_POINTS || m->wgt != nullptr, "lift: mesh modes need a map built by ivlm_lift_build_mesh");
    const long long total = (long long)B * m->n;
    lift_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        m->row_ptr, m->pix, m->wgt, masks, contact, B, m->V, m->n, (long long)m->H * m->W, mode, thr);
    h->launches++;
    IVLM_CHECK_CUDA(cudaGetLastError());
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
    ivlm_csr* m = new ivlm_csr();
    m->rows = rows; m->cols = cols; m->nnz = (long long)col.size();
    IVLM_CHECK_CUDA(cudaSetDevice(h->device));
    IVLM_TRY(upload(&m->row_ptr, row_ptr));
    IVLM_TRY(upload(&m->col, col));
    IVLM_TRY(upload(&m->val, val));
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
