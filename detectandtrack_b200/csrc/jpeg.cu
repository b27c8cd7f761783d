// Frame decode on the device (the reference reads frames with cv2.imread on loader threads,
// lib/utils/image.py:51-63): JPEG byte streams -> uint8 BGR frames [n, H, W, 3] in device memory through nvJPEG
// (NVJPEG_OUTPUT_BGRI = cv2.imread's channel order and interleaving), so only the COMPRESSED bytes cross PCIe and the frame
// lands where dt_prep_clip reads it.  nvJPEG is library code and sits outside the tensor hot path; it is dlopen'ed on first
// use, so libdt_b200.so loads (and everything else works) on a box without it.  One handle + decoder state per host thread
// (the loader threads decode in parallel, each on its own stream); Huffman decoding runs on the calling thread, the IDCT /
// colour conversion as GPU work on the caller's stream.
#include "common.cuh"
#include "../../include/dt_b200.h"
#include <dlfcn.h>
#include <nvjpeg.h>

namespace dt {

struct NvJpegApi {
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*);
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*);
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*);
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t);
  bool ok;
};

static const NvJpegApi& nvjpeg_api() {
  static const NvJpegApi api = [] {
    NvJpegApi a;
    memset(&a, 0, sizeof(a));
    void* h = nullptr;
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return a;
    a.CreateSimple = reinterpret_cast<decltype(a.CreateSimple)>(dlsym(h, "nvjpegCreateSimple"));
    a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(dlsym(h, "nvjpegJpegStateCreate"));
    a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(dlsym(h, "nvjpegGetImageInfo"));
    a.Decode = reinterpret_cast<decltype(a.Decode)>(dlsym(h, "nvjpegDecode"));
    a.ok = a.CreateSimple && a.JpegStateCreate && a.GetImageInfo && a.Decode;
    return a;
  }();
  return api;
}

struct ThreadDecoder { nvjpegHandle_t handle = nullptr; nvjpegJpegState_t state = nullptr; };

}  // namespace dt

using namespace dt;

extern "C" int dt_jpeg_available(void) { return nvjpeg_api().ok ? 1 : 0; }

extern "C" int dt_jpeg_decode(const unsigned char* const* jpegs, const size_t* sizes, int n, int H, int W, void* out_bgr, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DT_CHECK_ARG(n >= 0 && H >= 1 && W >= 1, "dt_jpeg_decode: bad shape n=%d H=%d W=%d", n, H, W);
  if (n == 0) return 0;
  DT_CHECK_ARG(jpegs && sizes && out_bgr, "dt_jpeg_decode: null pointer");
  const NvJpegApi& api = nvjpeg_api();
  DT_CHECK_ARG(api.ok, "dt_jpeg_decode: nvJPEG (libnvjpeg.so.12) is not available on this machine");
  static thread_local ThreadDecoder dec;
  if (!dec.handle) {
    DT_CHECK_ARG(api.CreateSimple(&dec.handle) == NVJPEG_STATUS_SUCCESS, "dt_jpeg_decode: nvjpegCreateSimple failed");
    DT_CHECK_ARG(api.JpegStateCreate(dec.handle, &dec.state) == NVJPEG_STATUS_SUCCESS, "dt_jpeg_decode: nvjpegJpegStateCreate failed");
  }
  unsigned char* out = static_cast<unsigned char*>(out_bgr);
  for (int i = 0; i < n; ++i) {
    DT_CHECK_ARG(jpegs[i] && sizes[i] > 0, "dt_jpeg_decode: image %d is empty", i);
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = api.GetImageInfo(dec.handle, jpegs[i], sizes[i], &nc, &ss, ws, hs);
    DT_CHECK_ARG(st == NVJPEG_STATUS_SUCCESS, "dt_jpeg_decode: image %d is not a JPEG stream nvJPEG can parse (status %d)", i, (int)st);
    DT_CHECK_ARG(ws[0] == W && hs[0] == H, "dt_jpeg_decode: image %d is %dx%d, expected %dx%d", i, hs[0], ws[0], H, W);
    nvjpegImage_t dst;
    memset(&dst, 0, sizeof(dst));
    dst.channel[0] = out + (size_t)i * H * W * 3;
    dst.pitch[0] = (size_t)W * 3;
    st = api.Decode(dec.handle, dec.state, jpegs[i], sizes[i], NVJPEG_OUTPUT_BGRI, &dst, stream);
    DT_CHECK_ARG(st == NVJPEG_STATUS_SUCCESS, "dt_jpeg_decode: nvjpegDecode failed on image %d (status %d)", i, (int)st);
  }
  return 0;
}
