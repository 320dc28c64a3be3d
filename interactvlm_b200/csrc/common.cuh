// Shared device helpers for the sm_100a kernels: PTX wrappers for mbarrier, TMA,
// tcgen05 (MMA / TMEM alloc / TMEM load / commit), plus small math utilities.
// Everything here is inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdio.h>
#include <stdint.h>

#define IVLM_DEVINL __device__ __forceinline__

namespace ivlm {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- misc
IVLM_DEVINL uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
IVLM_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

IVLM_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

IVLM_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
IVLM_DEVINL float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

IVLM_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
IVLM_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Activations. GELU is the exact erf form (torch.nn.GELU default); quick_gelu is
// x*sigmoid(1.702x) (HF CLIP "quick_gelu").
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICK_GELU = 2, ACT_RELU = 3, ACT_SILU = 4,
                 ACT_SWIGLU = 5 /* gemm_tcgen05 staged epilogue only: interleaved gate / up columns -> silu(gate) * up, N / 2 outputs */ };

IVLM_DEVINL float apply_act(float x, int act) {
    switch (act) {
        case ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        case ACT_QUICK_GELU: return x / (1.0f + __expf(-1.702f * x));
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_SILU: return x / (1.0f + __expf(-x));
        default: return x;
    }
}

// Raw SFU approximations (no range fix-up code: the arguments below are always well inside the valid range).
IVLM_DEVINL float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
IVLM_DEVINL float ex2_approx_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GELU (erf form) with the Abramowitz-Stegun 7.1.26 rational erf (|err| <= 1.5e-7, far below bf16 resolution):
// 0.5*x*(1+erf(x/sqrt2)); the negative branch uses 1+erf(-z) = poly*exp(-z^2) directly (no cancellation).
IVLM_DEVINL float gelu_fast(float x) {
    // with z = |x|/sqrt2: t = 1/(1 + p z), half_pe = 0.5 * poly(t) * t * exp(-z^2); constants pre-folded
    const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
    float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    poly = fmaf(t, poly, 0.5f * 1.421413741f);
    poly = fmaf(t, poly, 0.5f * -0.284496736f);
    poly = fmaf(t, poly, 0.5f * 0.254829592f);
    const float half_pe = poly * t * ex2_approx_ftz((x * x) * (-0.5f * 1.4426950408889634f));
    // x >= 0: x (1 - h) = x - |x| h;  x < 0: x h = -|x| h  ->  relu(x) - |x| h: one max and one fma instead of compare / select / mul
    return fmaf(-fabsf(x), half_pe, fmaxf(x, 0.f));
}
// two fp32 values rounded to bf16 (round-to-nearest-even) and widened again: ONE conversion instruction for the pair
// (cvt.rn.bf16x2.f32) plus two integer ops, instead of two conversions + two shifts -- the conversion unit is quarter-rate and
// shared with MUFU, which the activation epilogues of the GEMM already load
IVLM_DEVINL void bf16_round_pair(float& a, float& b) {
    const uint32_t p = pack_bf16x2(a, b);
    a = __uint_as_float(p << 16);
    b = __uint_as_float(p & 0xffff0000u);
}
// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 operations per issued instruction).  The activation epilogues
// and the softmax loops are instruction-issue-bound, not FMA-pipe-bound, so halving the instruction count of their fma chains
// is what speeds them up.  Same IEEE operations per element as the scalar forms (fma.rn / mul.rn / add.rn): identical results.
IVLM_DEVINL uint64_t f32x2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
IVLM_DEVINL void f32x2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
IVLM_DEVINL uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
IVLM_DEVINL uint64_t f32x2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
IVLM_DEVINL uint64_t f32x2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// gelu_fast on two values: the same operations in the same order per element (the polynomial carries the minus sign of
// relu(x) - |x| h in its constants), 9 issued instructions per element instead of 13
IVLM_DEVINL void gelu_fast_pair(float& x0, float& x1) {
    const uint64_t X = f32x2_pack(x0, x1);
    const uint64_t A = f32x2_pack(fabsf(x0), fabsf(x1));
    const float k = 0.3275911f * 0.70710678118654752440f;
    float a0, a1;
    f32x2_unpack(f32x2_fma(A, f32x2_pack(k, k), f32x2_pack(1.0f, 1.0f)), a0, a1);
    const uint64_t T = f32x2_pack(rcp_approx(a0), rcp_approx(a1));
    auto c2 = [](float c) { return f32x2_pack(c, c); };
    uint64_t P = f32x2_fma(T, c2(-0.5f * 1.061405429f), c2(-0.5f * -1.453152027f));
    P = f32x2_fma(T, P, c2(-0.5f * 1.421413741f));
    P = f32x2_fma(T, P, c2(-0.5f * -0.284496736f));
    P = f32x2_fma(T, P, c2(-0.5f * 0.254829592f));
    float e0, e1;
    f32x2_unpack(f32x2_mul(f32x2_mul(X, X), c2(-0.5f * 1.4426950408889634f)), e0, e1);
    const uint64_t NH = f32x2_mul(f32x2_mul(P, T), f32x2_pack(ex2_approx_ftz(e0), ex2_approx_ftz(e1)));   // -h
    f32x2_unpack(f32x2_fma(A, NH, f32x2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}
// act followed by the bf16 rounding the eager reference applies to the activation output
IVLM_DEVINL float apply_act_fast(float x, int act) {
    switch (act) {
        case ACT_GELU: return gelu_fast(x);
        case ACT_QUICK_GELU: return x * rcp_approx(1.0f + ex2_approx_ftz(x * (-1.702f * 1.4426950408889634f)));
        case ACT_RELU: return fmaxf(x, 0.0f);
        case ACT_SILU: return x * rcp_approx(1.0f + ex2_approx_ftz(x * -1.4426950408889634f));
        default: return x;
    }
}
// bf16x2 + bf16x2 with fp32 arithmetic and one rounding (torch's bf16 add)
IVLM_DEVINL uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    const float2 fa = unpack_bf16x2(a), fb = unpack_bf16x2(b);
    return pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
}
// One output sample of PyTorch's upsample_bilinear2d (align_corners=False) from src [.., sw] whose top-left (ch, cw) window is
// the source image: src coordinate = max((dst + 0.5) * scale - 0.5, 0), scale = in / out; same association as ATen:
// hy*(hx*a + lx*b) + ly*(hx*c + lx*d), every product and sum rounded separately.  Shared by bilinear_kernel and by the lift
// kernels that read the low-res logits directly, so both produce the same bits.
struct BilinearTap {
    int o00, o01, o10, o11;  // element offsets inside one [sh, sw] plane
    float hy, ly, hx, lx;
};
IVLM_DEVINL BilinearTap bilinear_tap(int y, int x, float sy, float sx, int ch, int cw, int sw) {
    const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f);
    const float fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < ch - 1 ? 1 : 0), x1 = x0 + (x0 < cw - 1 ? 1 : 0);
    BilinearTap t;
    t.ly = fy - (float)y0; t.lx = fx - (float)x0;
    t.hy = 1.f - t.ly; t.hx = 1.f - t.lx;
    t.o00 = y0 * sw + x0; t.o01 = y0 * sw + x1; t.o10 = y1 * sw + x0; t.o11 = y1 * sw + x1;
    return t;
}
IVLM_DEVINL float bilinear_combine(const BilinearTap& t, float v00, float v01, float v10, float v11) {
    const float top = __fadd_rn(__fmul_rn(t.hx, v00), __fmul_rn(t.lx, v01));
    const float bot = __fadd_rn(__fmul_rn(t.hx, v10), __fmul_rn(t.lx, v11));
    return __fadd_rn(__fmul_rn(t.hy, top), __fmul_rn(t.ly, bot));
}
IVLM_DEVINL float bilinear_eval(const BilinearTap& t, const float* __restrict__ s) {
    return bilinear_combine(t, __ldg(s + t.o00), __ldg(s + t.o01), __ldg(s + t.o10), __ldg(s + t.o11));
}
IVLM_DEVINL void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// (prologue, barrier / TMEM setup, prefetch of static operands) while its predecessor drains; it must execute this
// wait before touching anything the predecessor wrote.  Without the attribute the instruction returns immediately.
IVLM_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the NEXT kernel of the stream (if it was launched with the attribute) become resident now instead of when this grid
// has drained: its launch latency, prologue and static-operand prefetch then overlap this kernel's execution.  Takes
// effect once every CTA of this grid has executed it (or exited).  No-op without a programmatic dependent.
IVLM_DEVINL void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
IVLM_DEVINL void pdl_wait_then_launch() { pdl_wait(); pdl_launch(); }

// ---------------------------------------------------------------- mbarrier
IVLM_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
IVLM_DEVINL void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
IVLM_DEVINL void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
IVLM_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
IVLM_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
IVLM_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the thread (up to ~the hint, in ns) instead of spinning through
    // the issue slots the epilogue warps on the same SM sub-partition need; it wakes as soon as the phase completes
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
    return ok != 0;
}
IVLM_DEVINL uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug must trap (visible CUDA error) instead of hanging the GPU box.
IVLM_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0 && global_timer_ns() - t0 > 5000000000ull) {
            printf("ivlm: mbarrier wait timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- TMA
IVLM_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load: coordinates are (inner = element column, outer = row).
IVLM_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
IVLM_DEVINL void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_outer,
                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, "
        "%4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "l"(policy)
        : "memory");
}
// Pulls the box at (c_inner, c_outer) into L2 without a shared-memory destination or a barrier.
IVLM_DEVINL void tma_prefetch_2d_l2(const CUtensorMap* m, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c_inner),
                 "r"(c_outer)
                 : "memory");
}
IVLM_DEVINL uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
IVLM_DEVINL uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------- tcgen05
IVLM_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IVLM_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole-warp collectives (.sync.aligned): call with all 32 lanes converged.
IVLM_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
IVLM_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
IVLM_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
IVLM_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
IVLM_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
IVLM_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
IVLM_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory operand descriptor (sm_100 "version 1" format):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B), [46,48) version = 1,
//   [61,64) layout type (2 = SWIZZLE_128B).
IVLM_DEVINL uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) c_format (1 = F32), [7,10) a_format (1 = BF16), [10,13) b_format (1 = BF16),
//   [15] a_major (0 = K), [16] b_major (0 = K), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- legacy-path helpers (mma.sync / ldmatrix / cp.async)
IVLM_DEVINL void cp_async_16(void* smem_dst, const void* gsrc, bool pred) {
    uint32_t sz = pred ? 16u : 0u;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
                 : "memory");
}
IVLM_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
IVLM_DEVINL void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
IVLM_DEVINL void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(smem_row)));
}
IVLM_DEVINL void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(smem_row)));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
IVLM_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace ivlm
