// Context management and error reporting for the C ABI (include/ivlm_b200.h).
#include <stdarg.h>

#include "runtime.h"

namespace ivlm {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ivlm

extern "C" const char* ivlm_last_error(void) { return ivlm::g_err; }
extern "C" int ivlm_abi_version(void) { return 1; }

extern "C" int ivlm_create(ivlm_handle* out, int device) {
    IVLM_REQUIRE(out != nullptr, "ivlm_create: null out");
    int ndev = 0;
    IVLM_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    IVLM_REQUIRE(device >= 0 && device < ndev, "ivlm_create: device %d out of range (%d visible)", device, ndev);
    IVLM_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    IVLM_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    IVLM_REQUIRE(prop.major == 10, "ivlm_create: device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
    // the preprocessing / pose-refinement entry points take scratch from the stream-ordered allocator: keep what it has obtained
    // instead of handing it back to the driver at every synchronisation (the default release threshold is 0)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    ivlm_ctx* h = new ivlm_ctx();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return IVLM_OK;
}

extern "C" int ivlm_destroy(ivlm_handle h) {
    delete h;
    return IVLM_OK;
}

extern "C" uint64_t ivlm_launch_count(ivlm_handle h) { return h ? h->launches : 0; }

extern "C" int ivlm_set_option(ivlm_handle h, const char* name, int32_t value) {
    IVLM_REQUIRE(h && name, "set_option: null");
    if (std::string(name) == "pdl") {
        h->pdl = value ? 1 : 0;
        return IVLM_OK;
    }
    if (std::string(name) == "sm_limit") { h->sm_limit = value; return IVLM_OK; }
    if (std::string(name) == "gv_rows8_max_n") { h->gv_rows8_max_n = value; return IVLM_OK; }
    if (std::string(name) == "gv_warps") { h->gv_warps = value; return IVLM_OK; }
    if (std::string(name) == "gv_max_n") { h->gv_max_n = value; return IVLM_OK; }
    if (std::string(name) == "fused_split_force") { h->fused_split_force = value; return IVLM_OK; }
    if (std::string(name) == "gv_max_m") { h->gv_max_m = value; return IVLM_OK; }
    if (std::string(name) == "small_m_variant") {
        h->small_m_variant = value;
        return IVLM_OK;
    }
    if (std::string(name) == "global_attn_variant") {
        h->global_attn_variant = value;
        return IVLM_OK;
    }
    if (std::string(name) == "dec_prefetch") { h->dec_prefetch = value; return IVLM_OK; }
    if (std::string(name) == "ds_force_stream") { h->ds_force_stream = value; return IVLM_OK; }
    if (std::string(name) == "dec_warps") { h->dec_warps = value; return IVLM_OK; }
    if (std::string(name) == "ds_stages") { h->ds_stages = value; return IVLM_OK; }
    if (std::string(name) == "attn_variant") { h->attn_variant = value; return IVLM_OK; }
    if (std::string(name) == "attn_small_variant") { h->attn_small_variant = value; return IVLM_OK; }
    if (std::string(name) == "ds_prefetch_kb") { h->ds_prefetch_kb = value; return IVLM_OK; }
    if (std::string(name) == "attn_prefetch_ahead") { h->attn_prefetch_ahead = value; return IVLM_OK; }
    if (std::string(name) == "window_attn_variant") {
        h->window_attn_variant = value;
        return IVLM_OK;
    }
    ivlm::set_error("set_option: unknown option %s", name);
    return IVLM_ERR_ARG;
}

extern "C" int ivlm_set_workspace(ivlm_handle h, void* ptr, size_t bytes, void* stream) {
    IVLM_REQUIRE(h, "set_workspace: null handle");
    IVLM_REQUIRE(ptr == nullptr || bytes >= (1u << 20), "set_workspace: need at least 1 MiB");
    h->ws = reinterpret_cast<char*>(ptr);
    h->ws_bytes = ptr ? bytes : 0;
    if (ptr) IVLM_CHECK_CUDA(cudaMemsetAsync(ptr, 0, ivlm::IVLM_WS_COUNTER_BYTES, reinterpret_cast<cudaStream_t>(stream)));
    return IVLM_OK;
}
