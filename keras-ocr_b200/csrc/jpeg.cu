// jpeg.cu -- tools.read for JPEG input on the GPU.
//
// The reference decodes every input file on the host (tools.py:19-38: cv2.imread / cv2.imdecode + BGR->RGB) and the
// pipeline then uploads 3 bytes per pixel.  Here the compressed bytes are handed to nvJPEG (the CUDA toolkit's decoder --
// library code, like cuBLAS for a plain GEMM) which writes interleaved RGB straight into device memory, so only the
// compressed file crosses PCIe.  The library is opened lazily with dlopen: libb2ocr.so itself has no link-time
// dependency on it, and a box without it simply gets B2O_ERR_STATE from these two entry points (the Python wrapper then
// falls back to the host decoder for that file, as it does for PNG and for JPEG flavours nvJPEG refuses).
#include <dlfcn.h>
#include <nvjpeg.h>

#include "common.cuh"

namespace {

struct NvJpeg {
  void* lib = nullptr;
  nvjpegStatus_t (*create)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*state_create)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*state_destroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*info)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                           nvjpegImage_t*, cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};

template <typename F>
bool sym(void* lib, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(lib, name));
  return *out != nullptr;
}

NvJpeg* get(b2o_ctx* ctx) {
  if (ctx->jpeg) return static_cast<NvJpeg*>(ctx->jpeg);
  if (ctx->jpeg_failed) return nullptr;
  NvJpeg* j = new NvJpeg();
  for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
    j->lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (j->lib) break;
  }
  bool ok = j->lib && sym(j->lib, "nvjpegCreateSimple", &j->create) && sym(j->lib, "nvjpegDestroy", &j->destroy) &&
            sym(j->lib, "nvjpegJpegStateCreate", &j->state_create) && sym(j->lib, "nvjpegJpegStateDestroy", &j->state_destroy) &&
            sym(j->lib, "nvjpegGetImageInfo", &j->info) && sym(j->lib, "nvjpegDecode", &j->decode);
  ok = ok && j->create(&j->handle) == NVJPEG_STATUS_SUCCESS && j->state_create(j->handle, &j->state) == NVJPEG_STATUS_SUCCESS;
  if (!ok) {
    ctx->set_error(std::string("nvJPEG is not available: ") + (j->lib ? "initialisation failed" : "libnvjpeg.so.12 not found"));
    if (j->handle) j->destroy(j->handle);
    if (j->lib) dlclose(j->lib);
    delete j;
    ctx->jpeg_failed = true;
    return nullptr;
  }
  ctx->jpeg = j;
  return j;
}

}  // namespace

void jpeg_release(b2o_ctx* ctx) {
  NvJpeg* j = static_cast<NvJpeg*>(ctx->jpeg);
  if (!j) return;
  if (j->state) j->state_destroy(j->state);
  if (j->handle) j->destroy(j->handle);
  if (j->lib) dlclose(j->lib);
  delete j;
  ctx->jpeg = nullptr;
}

extern "C" int b2o_jpeg_info(b2o_ctx* ctx, const uint8_t* data, size_t size, int* height, int* width, int* components) {
  if (!ctx || !data || size < 4 || !height || !width || !components) return B2O_ERR_ARG;
  DeviceGuard guard(ctx->device);
  NvJpeg* j = get(ctx);
  if (!j) return B2O_ERR_STATE;
  int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t sub;
  if (j->info(j->handle, data, size, &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) {
    ctx->set_error("b2o_jpeg_info: not a JPEG stream nvJPEG can parse");
    return B2O_ERR_ARG;
  }
  *height = hs[0]; *width = ws[0]; *components = comps;
  return B2O_OK;
}

extern "C" int b2o_decode_jpeg(b2o_ctx* ctx, const uint8_t* data, size_t size, uint8_t* rgb_dev, int height, int width,
                               void* stream) {
  if (!ctx || !data || size < 4 || !rgb_dev || height <= 0 || width <= 0) return B2O_ERR_ARG;
  DeviceGuard guard(ctx->device);
  NvJpeg* j = get(ctx);
  if (!j) return B2O_ERR_STATE;
  nvjpegImage_t out;
  for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { out.channel[c] = nullptr; out.pitch[c] = 0; }
  out.channel[0] = rgb_dev;
  out.pitch[0] = static_cast<size_t>(width) * 3;              // interleaved RGB (cv2.imread + COLOR_BGR2RGB, tools.py:31-38)
  const nvjpegStatus_t st = j->decode(j->handle, j->state, data, size, NVJPEG_OUTPUT_RGBI, &out,
                                      reinterpret_cast<cudaStream_t>(stream));
  if (st != NVJPEG_STATUS_SUCCESS) {
    ctx->set_error("b2o_decode_jpeg: nvjpegDecode failed with status " + std::to_string(static_cast<int>(st)));
    return B2O_ERR_ARG;
  }
  ctx->launches++;                                            // nvJPEG's own kernels: counted as one library launch
  return B2O_OK;
}
