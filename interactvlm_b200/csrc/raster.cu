// "Render" of Render-Localise-Lift: mesh -> per-pixel (face, barycentrics) for V cameras, the pixel->vertex lift maps
// and the Phong-shaded SAM input views.  Replaces pytorch3d's MeshRasterizer / HardPhongShader as the reference drives
// them (preprocess_data/render_mesh_utils.py:115-198: blur_radius 0, faces_per_pixel 1, FoV perspective cameras,
// perspective-correct barycentrics, z_clip = znear / 2).
//
// Layout: projected vertices [V, Nv] float4 (x_ndc, y_ndc, z_view, -); 16x16-pixel tiles; per-tile face lists built by a
// count / scan / fill pass over face bounding boxes (list order is irrelevant: the depth test breaks ties on the face
// index, which equals pytorch3d's "first face in ascending order wins").  One CTA per (tile, view), one thread per pixel;
// face records are staged through shared memory in chunks.  All geometry arithmetic uses explicitly rounded fp32
// operations (no FMA contraction) in the operation order of pytorch3d's rasterize_meshes, so pix_to_face and the
// barycentrics are bit-identical to the CPU oracle (oracle/raster.py).
#include <vector>

#include "common.cuh"
#include "runtime.h"

namespace ivlm {

constexpr int RT = 16;            // tile side in pixels
constexpr int RCHUNK = 128;       // face records per shared-memory chunk
constexpr float R_EPS = 1e-8f;    // pytorch3d kEpsilon

struct DevCams {
    ivlm_raster_cam c[IVLM_RASTER_MAX_VIEWS];
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
// EdgeFunctionForward(p, a, b) = (p.x - a.x) * (b.y - a.y) - (p.y - a.y) * (b.x - a.x)
__device__ __forceinline__ float edge(float px, float py, float ax, float ay, float bx, float by) {
    return sub(mul(sub(px, ax), sub(by, ay)), mul(sub(py, ay), sub(bx, ax)));
}
// PixToNonSquareNdc on the flipped pixel index: -offset + (range i' + offset) / S1 with i' = S1 - 1 - i, range = 2 (S1/S2 when
// S1 is the longer side), offset = range / 2
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
    const float range = S1 > S2 ? mul(dvd((float)S1, (float)S2), 2.f) : 2.f;
    const float offset = dvd(range, 2.f);
    return add(-offset, dvd(add(mul(range, (float)(S1 - 1 - i)), offset), (float)S1));
}
__host__ __device__ __forceinline__ float ndc_half_range(int S1, int S2) { return S1 > S2 ? (float)S1 / (float)S2 : 1.f; }

__global__ void raster_project_kernel(const float* __restrict__ verts, int n_verts, DevCams cams, int V,
                                      float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (i >= n_verts) return;
    const ivlm_raster_cam& c = cams.c[v];
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        w[k] = add(add(add(mul(x, c.R[k]), mul(y, c.R[3 + k])), mul(z, c.R[6 + k])), c.T[k]);
    out[(size_t)v * n_verts + i] = make_float4(add(dvd(mul(w[0], c.fx), w[2]), c.cx), add(dvd(mul(w[1], c.fy), w[2]), c.cy), w[2], 0.f);
}

struct FaceBox {
    int tx0, tx1, ty0, ty1;  // inclusive tile range; tx0 > tx1 => culled
};

// status: 0 = rasterise, 1 = culled, 2 = skipped (straddles the camera plane)
// `pad` widens the box in NDC (sqrt(blur_radius) for the soft rasteriser, 0 otherwise)
__device__ __forceinline__ int face_box(const float4 a, const float4 b, const float4 c, float z_clip, int H, int W,
                                        FaceBox& box, float pad = 0.f) {
    box = {1, 0, 1, 0};
    const float zmin = fminf(a.z, fminf(b.z, c.z)), zmax = fmaxf(a.z, fmaxf(b.z, c.z));
    if (!(zmax >= fmaxf(z_clip, 0.f))) return 1;  // entirely behind the clip plane / the camera (or NaN)
    if (!(zmin > 0.f)) return 2;
    const float area = edge(c.x, c.y, a.x, a.y, b.x, b.y);
    if (area <= R_EPS && area >= -R_EPS) return 1;
    const float xmin = fminf(a.x, fminf(b.x, c.x)) - pad, xmax = fmaxf(a.x, fmaxf(b.x, c.x)) + pad;
    const float ymin = fminf(a.y, fminf(b.y, c.y)) - pad, ymax = fmaxf(a.y, fmaxf(b.y, c.y)) + pad;
    const float rx = ndc_half_range(W, H), ry = ndc_half_range(H, W);
    if (!(xmax >= -rx && xmin <= rx && ymax >= -ry && ymin <= ry)) return 1;
    // conservative pixel range (x decreases with the column, y with the row); the per-pixel test is exact
    const float fw = 0.5f * W / rx, fh = 0.5f * H / ry;
    const int j0 = (int)fmaxf(floorf((rx - fminf(xmax, rx)) * fw - 0.5f) - 1.f, 0.f);
    const int j1 = (int)fminf(ceilf((rx - fmaxf(xmin, -rx)) * fw - 0.5f) + 1.f, (float)(W - 1));
    const int i0 = (int)fmaxf(floorf((ry - fminf(ymax, ry)) * fh - 0.5f) - 1.f, 0.f);
    const int i1 = (int)fminf(ceilf((ry - fmaxf(ymin, -ry)) * fh - 0.5f) + 1.f, (float)(H - 1));
    if (j0 > j1 || i0 > i1) return 1;
    box = {j0 / RT, j1 / RT, i0 / RT, i1 / RT};
    return 0;
}

// pass 0: count faces per tile (+ skipped faces); pass 1: write face ids at tile_off[t] + cursor[t]++
template <int PASS>
__global__ void raster_bin_kernel(const float4* __restrict__ proj, const int* __restrict__ faces, int n_verts, int n_faces,
                                  DevCams cams, int H, int W, int tiles_x, int tiles_y, int* __restrict__ tile_cnt,
                                  const int* __restrict__ tile_off, int* __restrict__ tile_faces, int* __restrict__ n_skipped,
                                  float pad) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (f >= n_faces) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) return;
    const float4* p = proj + (size_t)v * n_verts;
    FaceBox box;
    const int st = face_box(p[i0], p[i1], p[i2], cams.c[v].z_clip, H, W, box, pad);
    if (st == 2 && PASS == 0) atomicAdd(n_skipped, 1);
    if (st != 0) return;
    for (int ty = box.ty0; ty <= box.ty1; ++ty)
        for (int tx = box.tx0; tx <= box.tx1; ++tx) {
            const int t = (v * tiles_y + ty) * tiles_x + tx;
            if (PASS == 0) {
                atomicAdd(tile_cnt + t, 1);
            } else {
                const int pos = atomicAdd(tile_cnt + t, 1);
                tile_faces[tile_off[t] + pos] = f;
            }
        }
}

// exclusive scan of n ints by one CTA; out[n] = total
__global__ void raster_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    const int b = t * per, e = min(b + per, n);
    int s = 0;
    for (int i = b; i < e; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < blockDim.x; d <<= 1) {
        const int x = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += x;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int i = b; i < e; ++i) {
        out[i] = run;
        run += in[i];
    }
    if (t == blockDim.x - 1) out[n] = part[t];
}

struct FaceRec {
    float x0, y0, x1, y1, x2, y2, z0, z1, z2, den;
    float xmin, xmax, ymin, ymax;
    int id, straddle;
};

__global__ void __launch_bounds__(RT* RT)
raster_tile_kernel(const float4* __restrict__ proj, const int* __restrict__ faces, int n_verts, DevCams cams, int H, int W,
                   int tiles_x, int tiles_y, const int* __restrict__ tile_off, const int* __restrict__ tile_faces,
                   int* __restrict__ pix_to_face, float* __restrict__ bary, float* __restrict__ zbuf,
                   long long* __restrict__ p2v) {
    __shared__ FaceRec rec[RCHUNK];
    const int v = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int t = (v * tiles_y + ty) * tiles_x + tx;
    const int lx = threadIdx.x % RT, ly = threadIdx.x / RT;
    const int col = tx * RT + lx, row = ty * RT + ly;
    const bool live = col < W && row < H;
    const float px = pix_to_ndc(col, W, H), py = pix_to_ndc(row, H, W);
    const float z_clip = cams.c[v].z_clip;
    const float4* p = proj + (size_t)v * n_verts;
    const int beg = tile_off[t], end = tile_off[t + 1];
    float best_z = INFINITY, b0 = -1.f, b1 = -1.f, b2 = -1.f;
    int best_f = -1;
    for (int base = beg; base < end; base += RCHUNK) {
        const int n = min(RCHUNK, end - base);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int f = tile_faces[base + k];
            const float4 a = p[faces[3 * f]], b = p[faces[3 * f + 1]], c = p[faces[3 * f + 2]];
            FaceRec r;
            r.x0 = a.x; r.y0 = a.y; r.z0 = a.z;
            r.x1 = b.x; r.y1 = b.y; r.z1 = b.z;
            r.x2 = c.x; r.y2 = c.y; r.z2 = c.z;
            r.den = add(edge(c.x, c.y, a.x, a.y, b.x, b.y), R_EPS);
            r.xmin = fminf(a.x, fminf(b.x, c.x)); r.xmax = fmaxf(a.x, fmaxf(b.x, c.x));
            r.ymin = fminf(a.y, fminf(b.y, c.y)); r.ymax = fmaxf(a.y, fmaxf(b.y, c.y));
            r.id = f;
            r.straddle = z_clip > 0.f && fminf(a.z, fminf(b.z, c.z)) < z_clip;
            rec[k] = r;
        }
        __syncthreads();
        if (!live) continue;
        for (int k = 0; k < n; ++k) {
            const FaceRec& r = rec[k];
            if (px > r.xmax || px < r.xmin || py > r.ymax || py < r.ymin) continue;  // CheckPointOutsideBoundingBox
            const float e0 = edge(px, py, r.x1, r.y1, r.x2, r.y2);
            const float e1 = edge(px, py, r.x2, r.y2, r.x0, r.y0);
            const float e2 = edge(px, py, r.x0, r.y0, r.x1, r.y1);
            // sign(b_k) = sign(e_k) * sign(den) (every later factor is positive): exact early rejection
            if (r.den > 0.f ? (e0 <= 0.f || e1 <= 0.f || e2 <= 0.f) : (e0 >= 0.f || e1 >= 0.f || e2 >= 0.f)) continue;
            const float w0 = dvd(e0, r.den), w1 = dvd(e1, r.den), w2 = dvd(e2, r.den);
            const float t0 = mul(mul(w0, r.z1), r.z2), t1 = mul(mul(r.z0, w1), r.z2), t2 = mul(mul(r.z0, r.z1), w2);
            const float d = fmaxf(add(add(t0, t1), t2), R_EPS);
            const float c0 = dvd(t0, d), c1 = dvd(t1, d), c2 = dvd(t2, d);
            const float pz = add(add(mul(c0, r.z0), mul(c1, r.z1)), mul(c2, r.z2));
            if (!(c0 > 0.f && c1 > 0.f && c2 > 0.f) || !(pz >= 0.f)) continue;
            if (r.straddle && !(pz >= z_clip)) continue;
            if (pz < best_z || (pz == best_z && r.id < best_f)) {
                best_z = pz; best_f = r.id; b0 = c0; b1 = c1; b2 = c2;
            }
        }
    }
    if (!live) return;
    const size_t o = ((size_t)v * H + row) * W + col;
    pix_to_face[o] = best_f;
    bary[3 * o] = b0; bary[3 * o + 1] = b1; bary[3 * o + 2] = b2;
    if (zbuf) zbuf[o] = best_f >= 0 ? best_z : -1.f;
    if (p2v) {
        p2v[3 * o] = best_f >= 0 ? faces[3 * best_f] : -1;
        p2v[3 * o + 1] = best_f >= 0 ? faces[3 * best_f + 1] : -1;
        p2v[3 * o + 2] = best_f >= 0 ? faces[3 * best_f + 2] : -1;
    }
}

// Meshes.verts_normals_packed(): per-corner cross products accumulated on the vertices (area weighting), normalised later
__global__ void raster_vertex_normals_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int n_faces,
                                             float* __restrict__ nrm) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const int i[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    float p[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) p[k][d] = verts[3 * i[k] + d];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = (k + 1) % 3, b = (k + 2) % 3;  // cross(v_a - v_k, v_b - v_k)
        const float ux = p[a][0] - p[k][0], uy = p[a][1] - p[k][1], uz = p[a][2] - p[k][2];
        const float vx = p[b][0] - p[k][0], vy = p[b][1] - p[k][1], vz = p[b][2] - p[k][2];
        atomicAdd(nrm + 3 * i[k], uy * vz - uz * vy);
        atomicAdd(nrm + 3 * i[k] + 1, uz * vx - ux * vz);
        atomicAdd(nrm + 3 * i[k] + 2, ux * vy - uy * vx);
    }
}

__device__ __forceinline__ void unit3(float& x, float& y, float& z) {
    const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-6f);
    x /= n; y /= n; z /= n;
}

// HardPhongShader with one PointLights and default Materials on top of the rasteriser output; white background;
// (rgb * 255) truncated to uint8 as render_mesh does (render_mesh_utils.py:195-197)
__global__ void raster_phong_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                                    const float* __restrict__ colors, const float* __restrict__ nrm, DevCams cams,
                                    const float* __restrict__ lights, int H, int W, const int* __restrict__ pix_to_face,
                                    const float* __restrict__ bary, float ka, float kd, float ks, float shininess,
                                    uint8_t* __restrict__ rgb) {
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (o >= (size_t)H * W) return;
    const size_t g = (size_t)v * H * W + o;
    const int f = pix_to_face[g];
    float out[3] = {1.f, 1.f, 1.f};
    if (f >= 0) {
        const float b[3] = {bary[3 * g], bary[3 * g + 1], bary[3 * g + 2]};
        float P[3] = {0, 0, 0}, N[3] = {0, 0, 0}, Cc[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = faces[3 * f + k];
            float nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
            unit3(nx, ny, nz);
            const float n3[3] = {nx, ny, nz};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                P[d] += b[k] * verts[3 * i + d];
                N[d] += b[k] * n3[d];
                Cc[d] += b[k] * colors[3 * i + d];
            }
        }
        unit3(N[0], N[1], N[2]);
        float L[3] = {lights[3 * v] - P[0], lights[3 * v + 1] - P[1], lights[3 * v + 2] - P[2]};
        unit3(L[0], L[1], L[2]);
        const float cosang = N[0] * L[0] + N[1] * L[1] + N[2] * L[2];
        const float diff = kd * fmaxf(cosang, 0.f);
        float E[3] = {cams.c[v].C[0] - P[0], cams.c[v].C[1] - P[1], cams.c[v].C[2] - P[2]};
        unit3(E[0], E[1], E[2]);
        float alpha = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) alpha += E[d] * (-L[d] + 2.f * cosang * N[d]);
        alpha = cosang > 0.f ? fmaxf(alpha, 0.f) : 0.f;
        const float spec = ks * powf(alpha, shininess);
#pragma unroll
        for (int d = 0; d < 3; ++d) out[d] = (ka + diff) * Cc[d] + spec;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) rgb[3 * g + d] = (uint8_t)(int)fminf(fmaxf(out[d] * 255.f, 0.f), 255.f);
}


// ------------------------------------------------------------------------------------------------ soft silhouette (optim/)
// SSRenderer (optim/renderer.py:64-104): MeshRasterizer(blur_radius, faces_per_pixel = K, perspective-correct, clipped
// barycentrics) + SoftSilhouetteShader(sigma).  Per pixel every face within sqrt(blur_radius) (squared NDC distance) is a
// fragment with a signed distance; the K fragments nearest in depth are kept (sorted insertion into [K][H*W] arrays, pixel
// index fastest so that neighbouring threads touch neighbouring words); alpha = 1 - prod(1 - sigmoid(-d / sigma)).

// pytorch3d PointLineDistanceForward: squared distance to the segment a-b and the clamped parameter t
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by, float& t) {
    const float dx = bx - ax, dy = by - ay;
    const float l2 = dx * dx + dy * dy;
    if (l2 <= R_EPS) {
        t = 1.f;
        return (px - bx) * (px - bx) + (py - by) * (py - by);
    }
    t = fminf(fmaxf((dx * (px - ax) + dy * (py - ay)) / l2, 0.f), 1.f);
    const float qx = ax + t * dx - px, qy = ay + t * dy - py;
    return qx * qx + qy * qy;
}

struct SoftFrag {
    float z, sd;
    int face;
};

// evaluates one face at one pixel; false when the face contributes no fragment.  `edge_t` returns the closest edge and its t.
__device__ __forceinline__ bool soft_eval(const float4 a, const float4 b, const float4 c, float px, float py, float blur,
                                          SoftFrag& out, int& edge_id, float& edge_t) {
    const float area = edge(c.x, c.y, a.x, a.y, b.x, b.y);
    if (area <= R_EPS && area >= -R_EPS) return false;
    const float den = area + R_EPS;
    const float w0 = edge(px, py, b.x, b.y, c.x, c.y) / den, w1 = edge(px, py, c.x, c.y, a.x, a.y) / den,
                w2 = edge(px, py, a.x, a.y, b.x, b.y) / den;
    const float t0 = w0 * b.z * c.z, t1 = a.z * w1 * c.z, t2 = a.z * b.z * w2;
    const float d = fmaxf(t0 + t1 + t2, R_EPS);
    const float b0 = t0 / d, b1 = t1 / d, b2 = t2 / d;
    float c0 = fminf(fmaxf(b0, 0.f), 1.f), c1 = fminf(fmaxf(b1, 0.f), 1.f), c2 = fminf(fmaxf(b2, 0.f), 1.f);
    const float cs = fmaxf(c0 + c1 + c2, 1e-5f);
    c0 /= cs; c1 /= cs; c2 /= cs;
    const float pz = c0 * a.z + c1 * b.z + c2 * c.z;
    if (!(pz >= 0.f)) return false;
    float ta, tb, tc;
    const float d0 = seg_dist2(px, py, a.x, a.y, b.x, b.y, ta), d1 = seg_dist2(px, py, b.x, b.y, c.x, c.y, tb),
                d2 = seg_dist2(px, py, c.x, c.y, a.x, a.y, tc);
    float dist = d0;
    edge_id = 0; edge_t = ta;
    if (d1 < dist) { dist = d1; edge_id = 1; edge_t = tb; }
    if (d2 < dist) { dist = d2; edge_id = 2; edge_t = tc; }
    const bool inside = b0 > 0.f && b1 > 0.f && b2 > 0.f;
    if (!inside && dist >= blur) return false;
    out.z = pz;
    out.sd = inside ? -dist : dist;
    return true;
}

__global__ void __launch_bounds__(RT* RT)
soft_tile_kernel(const float4* __restrict__ proj, const int* __restrict__ faces, int n_verts, int H, int W, int tiles_x,
                 const int* __restrict__ tile_off, const int* __restrict__ tile_faces, float sigma, float blur, int K,
                 float* __restrict__ alpha, float* __restrict__ zbuf0, int* __restrict__ n_frag, int* __restrict__ frag_face,
                 float* __restrict__ frag_sd, float* __restrict__ frag_z) {
    __shared__ float4 sa[RCHUNK], sb[RCHUNK], sc[RCHUNK];
    __shared__ int sid[RCHUNK];
    const int ty = blockIdx.y, tx = blockIdx.x;
    const int t = ty * tiles_x + tx;
    const int col = tx * RT + threadIdx.x % RT, row = ty * RT + threadIdx.x / RT;
    const bool live = col < W && row < H;
    const float px = pix_to_ndc(col, W, H), py = pix_to_ndc(row, H, W);
    const size_t hw = (size_t)H * W, o = (size_t)row * W + col;
    const int beg = tile_off[t], end = tile_off[t + 1];
    int n = 0;
    float zmin = INFINITY;
    for (int base = beg; base < end; base += RCHUNK) {
        const int cnt = min(RCHUNK, end - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int f = tile_faces[base + k];
            sa[k] = proj[faces[3 * f]]; sb[k] = proj[faces[3 * f + 1]]; sc[k] = proj[faces[3 * f + 2]];
            sid[k] = f;
        }
        __syncthreads();
        if (!live) continue;
        for (int k = 0; k < cnt; ++k) {
            const float4 a = sa[k], b = sb[k], c = sc[k];
            const float pad = sqrtf(blur);
            if (px > fmaxf(a.x, fmaxf(b.x, c.x)) + pad || px < fminf(a.x, fminf(b.x, c.x)) - pad ||
                py > fmaxf(a.y, fmaxf(b.y, c.y)) + pad || py < fminf(a.y, fminf(b.y, c.y)) - pad)
                continue;
            SoftFrag fr;
            int e;
            float et;
            if (!soft_eval(a, b, c, px, py, blur, fr, e, et)) continue;
            fr.face = sid[k];
            zmin = fminf(zmin, fr.z);
            // The K fragments nearest in (z, face) are kept in [k][pixel] global arrays, UNSORTED: alpha is a product and the
            // backward a sum over them, so only the set matters.  While the list is not full a fragment is appended (three
            // stores, no loads); once it is full (rare: K = 100 in the fit) the farthest entry is found by a scan and replaced
            // when the newcomer is nearer.  A sorted insertion costs ~n/2 dependent global round trips per fragment -- 2.4 ms
            // per forward at 256^2 with ~50 fragments per covered pixel.
            if (n < K) {
                frag_z[(size_t)n * hw + o] = fr.z;
                frag_face[(size_t)n * hw + o] = fr.face;
                frag_sd[(size_t)n * hw + o] = fr.sd;
                ++n;
            } else {
                int far = 0;
                float zf = frag_z[o];
                int ff = frag_face[o];
                for (int j = 1; j < K; ++j) {
                    const float zj = frag_z[(size_t)j * hw + o];
                    const int fj = frag_face[(size_t)j * hw + o];
                    if (zj > zf || (zj == zf && fj > ff)) { zf = zj; ff = fj; far = j; }
                }
                if (fr.z < zf || (fr.z == zf && fr.face < ff)) {
                    frag_z[(size_t)far * hw + o] = fr.z;
                    frag_face[(size_t)far * hw + o] = fr.face;
                    frag_sd[(size_t)far * hw + o] = fr.sd;
                }
            }
        }
    }
    if (!live) return;
    float keep = 1.f;
    for (int k = 0; k < n; ++k) keep *= 1.f / (1.f + __expf(-frag_sd[(size_t)k * hw + o] / sigma));  // 1 - sigmoid(-d/sigma)
    alpha[o] = 1.f - keep;
    zbuf0[o] = n > 0 ? zmin : -1.f;
    n_frag[o] = n;
}

__global__ void soft_backward_pixels_kernel(const float4* __restrict__ proj, const int* __restrict__ faces, int H, int W,
                                            float sigma, float blur, const float* __restrict__ grad_alpha,
                                            const int* __restrict__ n_frag, const int* __restrict__ frag_face,
                                            const float* __restrict__ frag_sd, float* __restrict__ g_ndc /*[Nv,2]*/) {
    const size_t hw = (size_t)H * W;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= hw) return;
    const float g = grad_alpha[o];
    const int n = n_frag[o];
    if (n == 0 || g == 0.f) return;
    const int row = (int)(o / W), col = (int)(o % W);
    const float px = pix_to_ndc(col, W, H), py = pix_to_ndc(row, H, W);
    float keep = 1.f;
    for (int k = 0; k < n; ++k) keep *= 1.f / (1.f + __expf(-frag_sd[(size_t)k * hw + o] / sigma));
    if (keep == 0.f) return;
    for (int k = 0; k < n; ++k) {
        const float sd = frag_sd[(size_t)k * hw + o];
        const float p = 1.f / (1.f + __expf(sd / sigma));
        // d alpha / d sd = -(prod_{j != k} (1 - p_j)) p_k (1 - p_k) / sigma = -keep p_k / sigma
        const float g_d = -g * keep * p / sigma * (sd < 0.f ? -1.f : 1.f);
        if (g_d == 0.f) continue;
        const int f = frag_face[(size_t)k * hw + o];
        const int iv[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
        const float4 a = proj[iv[0]], b = proj[iv[1]], c = proj[iv[2]];
        SoftFrag fr;
        int e;
        float t;
        if (!soft_eval(a, b, c, px, py, INFINITY, fr, e, t)) continue;
        const int ia = iv[e], ib = iv[(e + 1) % 3];
        const float4 va = e == 0 ? a : (e == 1 ? b : c), vb = e == 0 ? b : (e == 1 ? c : a);
        const float qx = va.x + t * (vb.x - va.x) - px, qy = va.y + t * (vb.y - va.y) - py;
        const float gx = g_d * 2.f * qx, gy = g_d * 2.f * qy;
        atomicAdd(g_ndc + 2 * ia, (1.f - t) * gx); atomicAdd(g_ndc + 2 * ia + 1, (1.f - t) * gy);
        atomicAdd(g_ndc + 2 * ib, t * gx); atomicAdd(g_ndc + 2 * ib + 1, t * gy);
    }
}

// NDC gradient -> view space (x_ndc = fx x / z + cx) -> world space (X_view = X_world R + T)
__global__ void soft_backward_verts_kernel(const float* __restrict__ verts, int n_verts, DevCams cams,
                                           const float* __restrict__ g_ndc, float* __restrict__ grad_verts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_verts) return;
    const ivlm_raster_cam& c = cams.c[0];
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = x * c.R[k] + y * c.R[3 + k] + z * c.R[6 + k] + c.T[k];
    const float gx = g_ndc[2 * i], gy = g_ndc[2 * i + 1];
    const float gv[3] = {gx * c.fx / w[2], gy * c.fy / w[2], -(gx * c.fx * w[0] + gy * c.fy * w[1]) / (w[2] * w[2])};
#pragma unroll
    for (int r = 0; r < 3; ++r) grad_verts[3 * i + r] = gv[0] * c.R[3 * r] + gv[1] * c.R[3 * r + 1] + gv[2] * c.R[3 * r + 2];
}

static int load_cams(const ivlm_raster_cam* cams_h, int V, DevCams& dc) {
    IVLM_REQUIRE(cams_h && V >= 1 && V <= IVLM_RASTER_MAX_VIEWS, "raster: need 1..%d cameras, got %d", IVLM_RASTER_MAX_VIEWS, V);
    for (int v = 0; v < V; ++v) dc.c[v] = cams_h[v];
    return IVLM_OK;
}

}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_rasterize_mesh(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                                   const ivlm_raster_cam* cams_h, int32_t V, int32_t H, int32_t W, int32_t* pix_to_face,
                                   float* bary, float* zbuf, int64_t* p2v, int32_t* n_skipped_h, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && verts && faces && pix_to_face && bary, "rasterize_mesh: null argument");
    IVLM_REQUIRE(n_verts > 0 && n_faces > 0 && H > 0 && W > 0, "rasterize_mesh: empty mesh or image (%d verts, %d faces, %dx%d)",
                 n_verts, n_faces, H, W);
    DevCams dc{};
    IVLM_TRY(load_cams(cams_h, V, dc));
    const int tiles_x = (W + RT - 1) / RT, tiles_y = (H + RT - 1) / RT;
    const int n_tiles = V * tiles_x * tiles_y;
    // one-time preprocessing call: scratch comes from the stream-ordered allocator and the tile-list size is read back once
    float4* proj = nullptr;
    int *tile_cnt = nullptr, *tile_off = nullptr, *tile_faces = nullptr, *skipped = nullptr;
    IVLM_CHECK_CUDA(cudaMallocAsync(&proj, sizeof(float4) * (size_t)V * n_verts, stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_cnt, sizeof(int) * (size_t)(n_tiles + 1), stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_off, sizeof(int) * (size_t)(n_tiles + 1), stream));
    skipped = tile_cnt + n_tiles;
    IVLM_CHECK_CUDA(cudaMemsetAsync(tile_cnt, 0, sizeof(int) * (size_t)(n_tiles + 1), stream));
    raster_project_kernel<<<dim3((n_verts + 255) / 256, V), 256, 0, stream>>>(verts, n_verts, dc, V, proj);
    const dim3 fgrid((n_faces + 127) / 128, V);
    raster_bin_kernel<0><<<fgrid, 128, 0, stream>>>(proj, faces, n_verts, n_faces, dc, H, W, tiles_x, tiles_y, tile_cnt,
                                                    nullptr, nullptr, skipped, 0.f);
    raster_scan_kernel<<<1, 1024, 0, stream>>>(tile_cnt, tile_off, n_tiles);
    int total = 0, n_skip = 0;
    IVLM_CHECK_CUDA(cudaMemcpyAsync(&total, tile_off + n_tiles, sizeof(int), cudaMemcpyDeviceToHost, stream));
    IVLM_CHECK_CUDA(cudaMemcpyAsync(&n_skip, skipped, sizeof(int), cudaMemcpyDeviceToHost, stream));
    IVLM_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (n_skipped_h) *n_skipped_h = n_skip;
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_faces, sizeof(int) * (size_t)(total > 0 ? total : 1), stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(tile_cnt, 0, sizeof(int) * (size_t)n_tiles, stream));
    raster_bin_kernel<1><<<fgrid, 128, 0, stream>>>(proj, faces, n_verts, n_faces, dc, H, W, tiles_x, tiles_y, tile_cnt,
                                                    tile_off, tile_faces, skipped, 0.f);
    raster_tile_kernel<<<dim3(tiles_x, tiles_y, V), RT * RT, 0, stream>>>(
        proj, faces, n_verts, dc, H, W, tiles_x, tiles_y, tile_off, tile_faces, pix_to_face, bary, zbuf,
        reinterpret_cast<long long*>(p2v));
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 6;
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_faces, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_off, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_cnt, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(proj, stream));
    return IVLM_OK;
}

extern "C" int ivlm_shade_phong(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                                const float* colors, const ivlm_raster_cam* cams_h, const float* lights_h, int32_t V, int32_t H,
                                int32_t W, const int32_t* pix_to_face, const float* bary, float ambient, float diffuse,
                                float specular, float shininess, uint8_t* rgb, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && verts && faces && colors && lights_h && pix_to_face && bary && rgb, "shade_phong: null argument");
    IVLM_REQUIRE(n_verts > 0 && n_faces > 0 && H > 0 && W > 0, "shade_phong: empty mesh or image");
    DevCams dc{};
    IVLM_TRY(load_cams(cams_h, V, dc));
    float *nrm = nullptr, *lights = nullptr;
    IVLM_CHECK_CUDA(cudaMallocAsync(&nrm, sizeof(float) * 3 * (size_t)n_verts, stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&lights, sizeof(float) * 3 * (size_t)V, stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(nrm, 0, sizeof(float) * 3 * (size_t)n_verts, stream));
    std::vector<float> lh(lights_h, lights_h + 3 * V);
    IVLM_CHECK_CUDA(cudaMemcpyAsync(lights, lh.data(), sizeof(float) * 3 * V, cudaMemcpyHostToDevice, stream));
    IVLM_CHECK_CUDA(cudaStreamSynchronize(stream));  // lh is a temporary
    raster_vertex_normals_kernel<<<(n_faces + 127) / 128, 128, 0, stream>>>(verts, faces, n_faces, nrm);
    raster_phong_kernel<<<dim3((unsigned)(((size_t)H * W + 255) / 256), V), 256, 0, stream>>>(
        verts, faces, colors, nrm, dc, lights, H, W, pix_to_face, bary, ambient, diffuse, specular, shininess, rgb);
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 2;
    IVLM_CHECK_CUDA(cudaFreeAsync(lights, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(nrm, stream));
    return IVLM_OK;
}

namespace ivlm {
// ------------------------------------------------------------------------------------------------ point-cloud rasteriser
// pytorch3d PointsRasterizer as preprocess_data/utils_obj_pc.py:88-113 reads it (`num_point2pixel == 1`): a pixel takes the
// point nearest in depth among those whose NDC distance to the pixel centre is below `radius` (z >= 0).  Every point splats
// a packed (depth bits, index) key into the pixels of its disc with atomicMin: positive floats order like their bit
// patterns, ties in depth go to the lower index -- deterministic whatever the execution order.
__global__ void points_splat_kernel(const float4* __restrict__ proj, int n_points, int H, int W, float radius,
                                    unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (i >= n_points) return;
    const float4 p = proj[(size_t)v * n_points + i];
    if (!(p.z >= 0.f)) return;
    const float rx = ndc_half_range(W, H), ry = ndc_half_range(H, W);
    if (!(p.x + radius >= -rx && p.x - radius <= rx && p.y + radius >= -ry && p.y - radius <= ry)) return;
    const float fw = 0.5f * W / rx, fh = 0.5f * H / ry;
    const int j0 = (int)fmaxf(floorf((rx - fminf(p.x + radius, rx)) * fw - 0.5f) - 1.f, 0.f);
    const int j1 = (int)fminf(ceilf((rx - fmaxf(p.x - radius, -rx)) * fw - 0.5f) + 1.f, (float)(W - 1));
    const int i0 = (int)fmaxf(floorf((ry - fminf(p.y + radius, ry)) * fh - 0.5f) - 1.f, 0.f);
    const int i1 = (int)fminf(ceilf((ry - fmaxf(p.y - radius, -ry)) * fh - 0.5f) + 1.f, (float)(H - 1));
    const float r2 = mul(radius, radius);
    const unsigned long long key = ((unsigned long long)__float_as_uint(p.z) << 32) | (unsigned int)i;
    for (int y = i0; y <= i1; ++y) {
        const float dy = sub(pix_to_ndc(y, H, W), p.y);
        for (int x = j0; x <= j1; ++x) {
            const float dx = sub(pix_to_ndc(x, W, H), p.x);
            if (add(mul(dx, dx), mul(dy, dy)) < r2) atomicMin(keys + ((size_t)v * H + y) * W + x, key);
        }
    }
}
__global__ void points_resolve_kernel(const unsigned long long* __restrict__ keys, long long n, long long* __restrict__ p2p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        p2p[i] = k == ~0ull ? -1 : (long long)(k & 0xffffffffull);
    }
}
}  // namespace ivlm

extern "C" int ivlm_rasterize_points(ivlm_handle h, const float* points, int32_t n_points, const ivlm_raster_cam* cams_h, int32_t V,
                                     int32_t H, int32_t W, float radius, int64_t* p2p, void* stream_) {
    using namespace ivlm;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && points && cams_h && p2p && n_points > 0 && V >= 1 && V <= IVLM_RASTER_MAX_VIEWS && H > 0 && W > 0 && radius > 0.f,
                 "rasterize_points: bad arguments (%d points, %d views, %dx%d, radius %g)", n_points, V, H, W, radius);
    DevCams dc{};
    IVLM_TRY(load_cams(cams_h, V, dc));
    float4* proj = nullptr;
    unsigned long long* keys = nullptr;
    const size_t npix = (size_t)V * H * W;
    IVLM_CHECK_CUDA(cudaMallocAsync(&proj, sizeof(float4) * (size_t)V * n_points, stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&keys, sizeof(unsigned long long) * npix, stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * npix, stream));
    raster_project_kernel<<<dim3((n_points + 255) / 256, V), 256, 0, stream>>>(points, n_points, dc, V, proj);
    points_splat_kernel<<<dim3((n_points + 127) / 128, V), 128, 0, stream>>>(proj, n_points, H, W, radius, keys);
    points_resolve_kernel<<<(unsigned)std::min<size_t>((npix + 255) / 256, 148 * 16), 256, 0, stream>>>(keys, (long long)npix,
                                                                                                      reinterpret_cast<long long*>(p2p));
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 3;
    IVLM_CHECK_CUDA(cudaFreeAsync(keys, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(proj, stream));
    return IVLM_OK;
}

extern "C" int ivlm_soft_silhouette(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts, int32_t n_faces,
                                    const ivlm_raster_cam* cam_h, int32_t H, int32_t W, float sigma, float blur_radius, int32_t K,
                                    float* alpha, float* zbuf0, int32_t* n_frag, int32_t* frag_face, float* frag_sd, float* frag_z,
                                    void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && verts && faces && alpha && zbuf0 && n_frag && frag_face && frag_sd && frag_z, "soft_silhouette: null argument");
    IVLM_REQUIRE(n_verts > 0 && n_faces > 0 && H > 0 && W > 0 && K >= 1 && sigma > 0.f && blur_radius >= 0.f,
                 "soft_silhouette: bad sizes (%d verts, %d faces, %dx%d, K=%d, sigma=%g)", n_verts, n_faces, H, W, K, sigma);
    DevCams dc{};
    IVLM_TRY(load_cams(cam_h, 1, dc));
    const int tiles_x = (W + RT - 1) / RT, tiles_y = (H + RT - 1) / RT, n_tiles = tiles_x * tiles_y;
    float4* proj = nullptr;
    int *tile_cnt = nullptr, *tile_off = nullptr, *tile_faces = nullptr;
    IVLM_CHECK_CUDA(cudaMallocAsync(&proj, sizeof(float4) * (size_t)n_verts, stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_cnt, sizeof(int) * (size_t)(n_tiles + 1), stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_off, sizeof(int) * (size_t)(n_tiles + 1), stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(tile_cnt, 0, sizeof(int) * (size_t)(n_tiles + 1), stream));
    const float pad = sqrtf(blur_radius);
    raster_project_kernel<<<dim3((n_verts + 255) / 256, 1), 256, 0, stream>>>(verts, n_verts, dc, 1, proj);
    const dim3 fgrid((n_faces + 127) / 128, 1);
    raster_bin_kernel<0><<<fgrid, 128, 0, stream>>>(proj, faces, n_verts, n_faces, dc, H, W, tiles_x, tiles_y, tile_cnt, nullptr,
                                                    nullptr, tile_cnt + n_tiles, pad);
    raster_scan_kernel<<<1, 1024, 0, stream>>>(tile_cnt, tile_off, n_tiles);
    // the per-tile face lists hold at most n_faces x n_tiles entries: when that bound is small (fit images are a few hundred
    // pixels wide and objects a few thousand faces) it is allocated outright from the stream-ordered pool and the iteration
    // has no host round trip; otherwise the exact total is read back first (one synchronisation)
    size_t cap = (size_t)n_faces * (size_t)n_tiles;
    if (cap > ((size_t)64 << 20)) {
        int total = 0;
        IVLM_CHECK_CUDA(cudaMemcpyAsync(&total, tile_off + n_tiles, sizeof(int), cudaMemcpyDeviceToHost, stream));
        IVLM_CHECK_CUDA(cudaStreamSynchronize(stream));
        cap = (size_t)(total > 0 ? total : 1);
    }
    IVLM_CHECK_CUDA(cudaMallocAsync(&tile_faces, sizeof(int) * cap, stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(tile_cnt, 0, sizeof(int) * (size_t)n_tiles, stream));
    raster_bin_kernel<1><<<fgrid, 128, 0, stream>>>(proj, faces, n_verts, n_faces, dc, H, W, tiles_x, tiles_y, tile_cnt, tile_off,
                                                    tile_faces, tile_cnt + n_tiles, pad);
    soft_tile_kernel<<<dim3(tiles_x, tiles_y), RT * RT, 0, stream>>>(proj, faces, n_verts, H, W, tiles_x, tile_off, tile_faces, sigma,
                                                                     blur_radius, K, alpha, zbuf0, n_frag, frag_face, frag_sd, frag_z);
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 6;
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_faces, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_off, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(tile_cnt, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(proj, stream));
    return IVLM_OK;
}

extern "C" int ivlm_soft_silhouette_backward(ivlm_handle h, const float* verts, const int32_t* faces, int32_t n_verts,
                                             int32_t n_faces, const ivlm_raster_cam* cam_h, int32_t H, int32_t W, float sigma,
                                             const float* grad_alpha, const int32_t* n_frag, const int32_t* frag_face,
                                             const float* frag_sd, float* grad_verts, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IVLM_REQUIRE(h && verts && faces && grad_alpha && n_frag && frag_face && frag_sd && grad_verts, "soft_silhouette_backward: null argument");
    IVLM_REQUIRE(n_verts > 0 && n_faces > 0 && H > 0 && W > 0 && sigma > 0.f, "soft_silhouette_backward: bad sizes");
    DevCams dc{};
    IVLM_TRY(load_cams(cam_h, 1, dc));
    float4* proj = nullptr;
    float* g_ndc = nullptr;
    IVLM_CHECK_CUDA(cudaMallocAsync(&proj, sizeof(float4) * (size_t)n_verts, stream));
    IVLM_CHECK_CUDA(cudaMallocAsync(&g_ndc, sizeof(float) * 2 * (size_t)n_verts, stream));
    IVLM_CHECK_CUDA(cudaMemsetAsync(g_ndc, 0, sizeof(float) * 2 * (size_t)n_verts, stream));
    raster_project_kernel<<<dim3((n_verts + 255) / 256, 1), 256, 0, stream>>>(verts, n_verts, dc, 1, proj);
    soft_backward_pixels_kernel<<<(unsigned)(((size_t)H * W + 127) / 128), 128, 0, stream>>>(proj, faces, H, W, sigma, 0.f, grad_alpha,
                                                                                            n_frag, frag_face, frag_sd, g_ndc);
    soft_backward_verts_kernel<<<(n_verts + 255) / 256, 256, 0, stream>>>(verts, n_verts, dc, g_ndc, grad_verts);
    IVLM_CHECK_CUDA(cudaGetLastError());
    h->launches += 3;
    IVLM_CHECK_CUDA(cudaFreeAsync(g_ndc, stream));
    IVLM_CHECK_CUDA(cudaFreeAsync(proj, stream));
    return IVLM_OK;
}
