// Library state, error reporting and device initialisation for libwfk_b200.so.
#include "internal.h"

namespace wfk {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};
int g_device = -1;
int g_num_sms = 0;
EncodeTiledFn g_encode_tiled = nullptr;
}  // namespace wfk

extern "C" const char* wfk_strerror(int status) {
  switch (status) {
    case WFK_OK: return "ok";
    case WFK_ERR_INVALID: return "invalid argument or unsupported shape";
    case WFK_ERR_CUDA: return "CUDA error";
    case WFK_ERR_NO_DEVICE: return "no sm_100 (B200) device";
    case WFK_ERR_NOT_INIT: return "wfk_init not called";
    default: return "unknown status";
  }
}

extern "C" const char* wfk_last_error(void) { return wfk::g_last_error; }
extern "C" int wfk_abi_version(void) { return WFK_ABI_VERSION; }
extern "C" int64_t wfk_launch_count(void) { return wfk::g_launches.load(); }

extern "C" int wfk_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return wfk::fail(WFK_ERR_NO_DEVICE, "cudaGetDeviceCount: %s (count=%d)", cudaGetErrorString(e), count);
  if (device < 0 || device >= count) return wfk::fail(WFK_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  WFK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return wfk::fail(WFK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major,
                     prop.minor);
  WFK_CUDA_CHECK(cudaSetDevice(device));
  WFK_CUDA_CHECK(cudaFree(0));
  if (wfk::g_encode_tiled == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    WFK_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return wfk::fail(WFK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    wfk::g_encode_tiled = reinterpret_cast<wfk::EncodeTiledFn>(fn);
  }
  wfk::g_num_sms = prop.multiProcessorCount;
  wfk::g_device = device;
  return WFK_OK;
}
