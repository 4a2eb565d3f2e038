// Library state, error reporting and device initialisation for libwfk_b200.so.
#include "internal.h"

namespace wfk {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<uint64_t> g_init_mask{0};
int g_num_sms_dev[kMaxDevices] = {0};
thread_local int t_device = -1;
int* g_flag_host[kMaxDevices] = {nullptr};
int* g_flag_dev[kMaxDevices] = {nullptr};
EncodeTiledFn g_encode_tiled = nullptr;
static std::mutex g_init_mu;

void DeviceScope::enter(int device) {
  dev = device;
  saved = t_device;
  if (device < 0 || device >= kMaxDevices || !(g_init_mask.load(std::memory_order_acquire) & (1ull << device))) {
    status = fail(WFK_ERR_NOT_INIT, "wfk_init(%d) was not called for the device this call targets", device);
    dev = -2;
    return;
  }
  cudaError_t e = cudaGetDevice(&prev);
  if (e == cudaSuccess && prev != device) {
    e = cudaSetDevice(device);
    switched = (e == cudaSuccess);
  }
  if (e != cudaSuccess) {
    status = fail(WFK_ERR_CUDA, "selecting device %d failed: %s", device, cudaGetErrorString(e));
    dev = -2;
    return;
  }
  t_device = device;
}

DeviceScope DeviceScope::from_stream(void* stream) {
  int device = -1;
  cudaError_t e;
  // cudaStreamGetDevice is illegal on a stream that is being captured into a CUDA graph (it invalidates the capture,
  // probed in scripts/probes/capture_api_probe.cu): a capturing thread's current device is the stream's device
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (stream != nullptr && cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cap) == cudaSuccess &&
      cap != cudaStreamCaptureStatusNone) {
    e = cudaGetDevice(&device);
  } else {
    // the legacy / per-thread default streams report the calling thread's current device
    e = cudaStreamGetDevice(static_cast<cudaStream_t>(stream), &device);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaGetDevice(&device);
    }
  }
  if (e != cudaSuccess) device = -1;
  return DeviceScope(device);
}

DeviceScope DeviceScope::from_stream_or_pointer(void* stream, const void* device_ptr) {
  // the pointer decides (capture-safe, unambiguous); the stream only when the call has no usable device pointer
  cudaPointerAttributes attr;
  if (device_ptr != nullptr && cudaPointerGetAttributes(&attr, device_ptr) == cudaSuccess &&
      (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
    return DeviceScope(attr.device);
  cudaGetLastError();
  return from_stream(stream);
}

DeviceScope DeviceScope::from_pointer(const void* device_ptr) {
  int device = -1;
  cudaPointerAttributes attr;
  if (device_ptr != nullptr && cudaPointerGetAttributes(&attr, device_ptr) == cudaSuccess &&
      (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
    device = attr.device;
  else
    cudaGetLastError();
  if (device < 0) cudaGetDevice(&device);
  return DeviceScope(device);
}

DeviceScope::~DeviceScope() {
  if (dev == -2) return;  // never entered (or moved from)
  t_device = saved;
  if (switched) cudaSetDevice(prev);
}
}  // namespace wfk

extern "C" const char* wfk_strerror(int status) {
  switch (status) {
    case WFK_OK: return "ok";
    case WFK_ERR_INVALID: return "invalid argument or unsupported shape";
    case WFK_ERR_CUDA: return "CUDA error";
    case WFK_ERR_NO_DEVICE: return "no sm_100 (B200) device";
    case WFK_ERR_NOT_INIT: return "wfk_init not called";
    case WFK_ERR_NONFINITE: return "non-finite activation (fp16 range exceeded)";
    default: return "unknown status";
  }
}

extern "C" const char* wfk_last_error(void) { return wfk::g_last_error; }
extern "C" int wfk_abi_version(void) { return WFK_ABI_VERSION; }
extern "C" int64_t wfk_launch_count(void) { return wfk::g_launches.load(); }

extern "C" int wfk_nonfinite_status(int device, int reset) {
  if (device < 0 || device >= wfk::kMaxDevices || wfk::g_flag_host[device] == nullptr)
    return wfk::fail(WFK_ERR_NOT_INIT, "wfk_init(%d) was not called", device);
  volatile int* f = wfk::g_flag_host[device];
  const int v = *f;
  if (v != 0 && reset) *f = 0;
  if (v != 0)
    return wfk::fail(WFK_ERR_NONFINITE,
                     "a GroupNorm statistic or model output went inf / NaN on device %d (stage mask 0x%x): with fp16 "
                     "activations this is what a value beyond 65504 turns into; bf16 operands have the fp32 range", device, v);
  return WFK_OK;
}

// Registers `device` with the library (any number of devices, from any thread). The caller's current device is left
// as it was: every later entry point selects the device of its stream / pointers itself.
extern "C" int wfk_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return wfk::fail(WFK_ERR_NO_DEVICE, "cudaGetDeviceCount: %s (count=%d)", cudaGetErrorString(e), count);
  if (device < 0 || device >= count || device >= wfk::kMaxDevices)
    return wfk::fail(WFK_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  std::lock_guard<std::mutex> lock(wfk::g_init_mu);
  if (wfk::g_init_mask.load() & (1ull << device)) return WFK_OK;
  cudaDeviceProp prop;
  WFK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return wfk::fail(WFK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major,
                     prop.minor);
  int prev = -1;
  WFK_CUDA_CHECK(cudaGetDevice(&prev));
  WFK_CUDA_CHECK(cudaSetDevice(device));
  cudaError_t ce = cudaFree(0);   // creates the primary context
  if (ce == cudaSuccess && wfk::g_encode_tiled == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ce == cudaSuccess && (fn == nullptr || qres != cudaDriverEntryPointSuccess)) {
      cudaSetDevice(prev);
      return wfk::fail(WFK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    }
    if (ce == cudaSuccess) wfk::g_encode_tiled = reinterpret_cast<wfk::EncodeTiledFn>(fn);
  }
  if (ce == cudaSuccess && wfk::g_flag_host[device] == nullptr) {
    int* hp = nullptr;
    int* dp = nullptr;
    ce = cudaHostAlloc(reinterpret_cast<void**>(&hp), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
    if (ce == cudaSuccess) {
      *hp = 0;
      ce = cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), hp, 0);
    }
    if (ce == cudaSuccess) {
      wfk::g_flag_host[device] = hp;
      wfk::g_flag_dev[device] = dp;
    }
  }
  if (prev != device) cudaSetDevice(prev);
  if (ce != cudaSuccess) return wfk::fail(WFK_ERR_CUDA, "initialising device %d failed: %s", device, cudaGetErrorString(ce));
  wfk::g_num_sms_dev[device] = prop.multiProcessorCount;
  wfk::g_init_mask.fetch_or(1ull << device, std::memory_order_release);
  return WFK_OK;
}
