// DRAFT / TESTBED -- NOT part of libivlm_b200.so and NOT validated on hardware yet (written at the end of round 1 with the
// GPU budget spent; it compiles for sm_100a).  Purpose: bring up the 2-CTA (cta_group::2) form of the tcgen05 GEMM that
// DESIGN.md section 8 lists as the next step, outside the product library, with its own correctness check and timing:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I interactvlm_b200/csrc -I include \
//        tools/drafts/gemm_2cta_testbed.cu -o gpurun_out/gemm_2cta -lcuda
//   timeout 60 gpurun_out/gemm_2cta            # every wait is bounded (trap after 5 s), wrap in `timeout` anyway
//
// out[M,N] = A[M,K] . W[N,K]^T, bf16 in, fp32 accumulate, bf16 out.  One CTA PAIR (cluster of 2, same TPC) per 256 x 256
// tile: CTA r of the pair holds rows [128 r, 128 r + 128) of the A tile and rows [128 r, 128 r + 128) of the W tile in its
// shared memory (32 KB per stage instead of 48 KB for the 1-CTA 128 x 256 tile), the leader's single thread issues
// tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16) which reads both CTAs' operand halves, and each CTA's TMEM receives the
// 128 accumulator rows it owns.  Protocol (after DeepGEMM / CUTLASS SM100 2-SM kernels):
//   full[s]   lives in the LEADER: it expects the bytes of both CTAs' TMA loads (the peer's loads name the leader's barrier),
//   empty[s]  lives in both CTAs, released by the leader's tcgen05.commit ... multicast::cluster (mask 0b11),
//   tfull[a]  lives in both CTAs (same multicast commit after the last k-block of a tile),
//   tempty[a] lives in the LEADER: the epilogue warps of both CTAs arrive on it (remote arrive through mapa).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

using namespace ivlm;

constexpr int T_BM = 128, T_BN_HALF = 128, T_BN = 256, T_BK = 64, T_STAGES = 6;
constexpr int T_A_BYTES = T_BM * T_BK * 2, T_B_BYTES = T_BN_HALF * T_BK * 2, T_STAGE = T_A_BYTES + T_B_BYTES;
constexpr int T_EPI_WARP0 = 4, T_EPI_WARPS = 8, T_THREADS = (T_EPI_WARP0 + T_EPI_WARPS) * 32;
constexpr int T_SMEM = T_STAGES * T_STAGE + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++spins & 255u) == 0 && global_timer_ns() - t0 > 5000000000ull) {
            printf("2cta: mbarrier wait timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// TMA load of this CTA's operand half; the transaction bytes are credited to `mbar_cluster_addr` (the LEADER's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c_inner,
                                                int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T_THREADS, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* __restrict__ out,
                 int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + T_STAGES * T_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_STAGES * T_STAGE);
    uint64_t *full_bar = bars, *empty_bar = bars + T_STAGES, *tfull_bar = bars + 2 * T_STAGES, *tempty_bar = bars + 2 * T_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_m = M / (2 * T_BM), tiles_n = N / T_BN, n_tiles = tiles_m * tiles_n, k_blocks = K / T_BK;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);   // used in the leader only: one arrive.expect_tx, bytes from both CTAs
            mbar_init(&empty_bar[s], 1);  // one multicast commit per phase
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 2 * T_EPI_WARPS);  // used in the leader only: the epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ------------------------------------------------ TMA producer (both CTAs)
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < n_tiles; t += n_pairs) {
                const int m0 = (t / tiles_n) * (2 * T_BM) + (int)rank * T_BM, n0 = (t % tiles_n) * T_BN + (int)rank * T_BN_HALF;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * T_STAGE);
                    const uint32_t leader_full = map_to_rank(smem_u32(&full_bar[stage]), 0);
                    tma_load_2d_2sm(smem_a + stage * T_A_BYTES, &tmA, leader_full, kb * T_BK, m0);
                    tma_load_2d_2sm(smem_b + stage * T_B_BYTES, &tmB, leader_full, kb * T_BK, n0);
                    if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {  // ---------------------------------------- MMA issuer (leader only)
            constexpr uint32_t idesc = umma_idesc_bf16(2 * T_BM, T_BN);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
                const int as = it & 1;
                mbar_wait_cluster(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * T_BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_cluster(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * T_A_BYTES));
                    const uint64_t db = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * T_B_BYTES));
#pragma unroll
                    for (int k = 0; k < T_BK / 16; ++k) umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tfull_bar[as]);
            }
        }
    } else if (warp >= T_EPI_WARP0) {  // ------------------------------------ epilogue (both CTAs, own 128 rows)
        const int quad = warp & 3, chalf = (warp - T_EPI_WARP0) >> 2;
        int it = 0;
        for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
            const int as = it & 1;
            const int m0 = (t / tiles_n) * (2 * T_BM) + (int)rank * T_BM, n0 = (t % tiles_n) * T_BN;
            mbar_wait_cluster(&tfull_bar[as], (it >> 1) & 1);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
#pragma unroll 1
            for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(as * T_BN + c0), raw);
                tmem_ld_wait();
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * N + n0 + c0);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint4 q;
                    q.x = pack_bf16x2(__uint_as_float(raw[j8 * 8 + 0]), __uint_as_float(raw[j8 * 8 + 1]));
                    q.y = pack_bf16x2(__uint_as_float(raw[j8 * 8 + 2]), __uint_as_float(raw[j8 * 8 + 3]));
                    q.z = pack_bf16x2(__uint_as_float(raw[j8 * 8 + 4]), __uint_as_float(raw[j8 * 8 + 5]));
                    q.w = pack_bf16x2(__uint_as_float(raw[j8 * 8 + 6]), __uint_as_float(raw[j8 * 8 + 7]));
                    dst[j8] = q;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(map_to_rank(smem_u32(&tempty_bar[as]), 0));
        }
    }
    tc_fence_before();
    cluster_sync_all();  // nobody frees TMEM or exits while the peer may still read / signal
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

__global__ void ref_gemm_kernel(const __nv_bfloat16* a, const __nv_bfloat16* w, float* out, int M, int N, int K, int rows, int cols) {
    const int r = blockIdx.y * (M / rows), c = (blockIdx.x * blockDim.x + threadIdx.x) * (N / cols);  // a sparse sample of outputs
    if (blockIdx.x * blockDim.x + threadIdx.x >= cols) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(a[(size_t)r * K + k]) * __bfloat162float(w[(size_t)c * K + k]);
    out[blockIdx.y * cols + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static CUtensorMap make_map(void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    const int M = 32768, N = 5120, K = 1280;  // SAM MLP-1 at 8 views
    std::vector<__nv_bfloat16> ha((size_t)M * K), hw((size_t)N * K);
    srand(1);
    for (auto& x : ha) x = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    for (auto& x : hw) x = __float2bfloat16((rand() % 2001 - 1000) / 16000.f);
    __nv_bfloat16 *a, *w, *out;
    float* ref;
    CK(cudaMalloc(&a, ha.size() * 2)); CK(cudaMalloc(&w, hw.size() * 2)); CK(cudaMalloc(&out, (size_t)M * N * 2));
    const int RS = 64, CS = 256;
    CK(cudaMalloc(&ref, RS * CS * 4));
    CK(cudaMemcpy(a, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(out, 0, (size_t)M * N * 2));
    const CUtensorMap ta = make_map(a, M, K, T_BM), tb = make_map(w, N, K, T_BN_HALF);
    CK(cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = (sms / 2) * 2;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gemm_2cta_kernel<<<grid, T_THREADS, T_SMEM>>>(ta, tb, out, M, N, K);
    CK(cudaDeviceSynchronize());
    ref_gemm_kernel<<<dim3((CS + 127) / 128, RS), 128>>>(a, w, ref, M, N, K, RS, CS);
    CK(cudaDeviceSynchronize());
    std::vector<float> hr(RS * CS);
    std::vector<__nv_bfloat16> ho((size_t)M * N);
    CK(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    for (int i = 0; i < RS; ++i)
        for (int j = 0; j < CS; ++j) {
            const float got = __bfloat162float(ho[(size_t)(i * (M / RS)) * N + j * (N / CS)]);
            max_err = fmax(max_err, fabs(got - hr[i * CS + j]));
            max_ref = fmax(max_ref, fabs(hr[i * CS + j]));
        }
    printf("max abs err %.4g on scale %.4g (%s)\n", max_err, max_ref, max_err < 0.02 * max_ref ? "OK" : "MISMATCH");
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) gemm_2cta_kernel<<<grid, T_THREADS, T_SMEM>>>(ta, tb, out, M, N, K);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("2-CTA 256x256 tile, M=%d N=%d K=%d: %.4f ms  %.1f TFLOP/s (1-CTA kernel with bias+GELU on this shape: 1208)\n", M, N, K,
           ms / reps, 2.0 * M * N * K / (ms / reps) / 1e9);
    return 0;
}
