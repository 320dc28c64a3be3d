// JPEG decode on the GPU (input pipeline so that 8 GPUs are not starved by 8 host cores).  The reference decodes
// on the CPU with cv2.imread (run_demo.py:330, datasets/*): here the compressed bytes go to nvJPEG (a CUDA library call, like
// cuBLAS for a plain GEMM) and the pixels land in device memory as interleaved RGB, ready for ivlm_resample_u8 /
// ivlm_preprocess_u8_bf16.  nvJPEG's IDCT / chroma upsampling can differ from libjpeg-turbo by a few grey levels: the decode is
// validated against OpenCV with a tolerance, not bit for bit (tests/test_harness_gpu.py).
#include <nvjpeg.h>

#include "runtime.h"

namespace ivlm {
struct JpegState {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
};
static int jpeg_state(ivlm_ctx* h, JpegState** out) {
    if (h->jpeg == nullptr) {
        JpegState* js = new JpegState();
        if (nvjpegCreateSimple(&js->handle) != NVJPEG_STATUS_SUCCESS || nvjpegJpegStateCreate(js->handle, &js->state) != NVJPEG_STATUS_SUCCESS) {
            delete js;
            set_error("jpeg: nvjpegCreateSimple / nvjpegJpegStateCreate failed");
            return IVLM_ERR_CUDA;
        }
        h->jpeg = js;
    }
    *out = reinterpret_cast<JpegState*>(h->jpeg);
    return IVLM_OK;
}
}  // namespace ivlm

using namespace ivlm;

extern "C" int ivlm_jpeg_info(ivlm_handle h, const uint8_t* data_h, size_t n, int32_t* height, int32_t* width) {
    IVLM_REQUIRE(h && data_h && n > 0 && height && width, "jpeg_info: bad arguments");
    JpegState* js;
    IVLM_TRY(jpeg_state(h, &js));
    int comps, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t sub;
    const nvjpegStatus_t st = nvjpegGetImageInfo(js->handle, data_h, n, &comps, &sub, ws, hs);
    IVLM_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "jpeg_info: not a decodable JPEG stream (nvjpeg status %d)", (int)st);
    *width = ws[0];
    *height = hs[0];
    return IVLM_OK;
}

// data_h: the compressed file in HOST memory; rgb: DEVICE [H,W,3] uint8 (H, W from ivlm_jpeg_info).  Enqueues on `stream`.
extern "C" int ivlm_jpeg_decode_rgb(ivlm_handle h, const uint8_t* data_h, size_t n, uint8_t* rgb, int32_t H, int32_t W, void* stream) {
    IVLM_REQUIRE(h && data_h && n > 0 && rgb && H > 0 && W > 0, "jpeg_decode: bad arguments");
    JpegState* js;
    IVLM_TRY(jpeg_state(h, &js));
    nvjpegImage_t img = {};
    img.channel[0] = rgb;
    img.pitch[0] = (size_t)W * 3;
    const nvjpegStatus_t st = nvjpegDecode(js->handle, js->state, data_h, n, NVJPEG_OUTPUT_RGBI, &img, reinterpret_cast<cudaStream_t>(stream));
    IVLM_REQUIRE(st == NVJPEG_STATUS_SUCCESS, "jpeg_decode: nvjpegDecode failed (status %d)", (int)st);
    h->launches++;
    return IVLM_OK;
}
