// Internal helpers shared by the translation units of libwfk_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/wfk_b200.h"

namespace wfk {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;
// Per-device state: one process may drive several B200s (and several host threads): nothing below is "the" device.
constexpr int kMaxDevices = 64;
extern std::atomic<uint64_t> g_init_mask;     // bit d set once wfk_init(d) succeeded
extern int g_num_sms_dev[kMaxDevices];
extern thread_local int t_device;             // device of the innermost DeviceScope of this thread (-1 outside)
inline int num_sms() { return t_device >= 0 ? g_num_sms_dev[t_device] : 0; }
// Non-finite guard: one host-mapped int per device, set (never cleared) by the kernels that see GroupNorm statistics or
// model outputs go inf / NaN -- what an fp16 activation beyond 65504 turns into one layer later. Read through
// wfk_nonfinite_status() without any CUDA call; definitive after the stream has been synchronised.
extern int* g_flag_host[kMaxDevices];
extern int* g_flag_dev[kMaxDevices];
inline int* nonfinite_flag() { return t_device >= 0 ? g_flag_dev[t_device] : nullptr; }

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

#define WFK_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::wfk::fail(WFK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                               \
  } while (0)

#define WFK_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::wfk::fail(WFK_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Every C-ABI entry point runs inside a DeviceScope: the device is taken from the caller's stream (or, for calls
// without a stream, from a device pointer), must have been initialised with wfk_init, is made current for the
// duration of the call and the caller's current device is restored on return.
struct DeviceScope {
  int status = WFK_OK;
  int dev = -1, prev = -1, saved = -1;
  bool switched = false;
  void enter(int device);
  explicit DeviceScope(int device) { enter(device); }
  static DeviceScope from_stream(void* stream);
  static DeviceScope from_stream_or_pointer(void* stream, const void* device_ptr);
  static DeviceScope from_pointer(const void* device_ptr);
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope(DeviceScope&& o) noexcept : status(o.status), dev(o.dev), prev(o.prev), saved(o.saved), switched(o.switched) {
    o.switched = false;
    o.dev = -2;
  }
  ~DeviceScope();
};
#define WFK_ENTER_STREAM(stream)                                            \
  ::wfk::DeviceScope _wfk_scope = ::wfk::DeviceScope::from_stream(stream);  \
  if (_wfk_scope.status != WFK_OK) return _wfk_scope.status
// The usual form: the device of `devptr` (a device pointer argument of the call). A stream handle alone cannot tell:
// the legacy default stream is handle 0 on EVERY device, which is what PyTorch hands over for a device's default stream.
#define WFK_ENTER(stream, devptr)                                                            \
  ::wfk::DeviceScope _wfk_scope = ::wfk::DeviceScope::from_stream_or_pointer(stream, devptr); \
  if (_wfk_scope.status != WFK_OK) return _wfk_scope.status
#define WFK_ENTER_DEVICE(dev)             \
  ::wfk::DeviceScope _wfk_scope(dev);     \
  if (_wfk_scope.status != WFK_OK) return _wfk_scope.status
#define WFK_ENTER_PTR(ptr)                                                  \
  ::wfk::DeviceScope _wfk_scope = ::wfk::DeviceScope::from_pointer(ptr);    \
  if (_wfk_scope.status != WFK_OK) return _wfk_scope.status

// "Once per device" guard for cudaFuncSetAttribute and similar per-device, per-function settings.
struct PerDeviceOnce {
  std::mutex mu;
  std::atomic<uint64_t> done{0};
  struct Lock {
    PerDeviceOnce& o;
    bool need = false, locked = false;
    explicit Lock(PerDeviceOnce& once) : o(once) {
      const uint64_t bit = 1ull << (t_device < 0 ? 0 : t_device);
      if (o.done.load(std::memory_order_acquire) & bit) return;
      o.mu.lock();
      locked = true;
      need = !(o.done.load(std::memory_order_relaxed) & bit);
    }
    bool needed() const { return need; }
    void finished() { o.done.fetch_or(1ull << (t_device < 0 ? 0 : t_device), std::memory_order_release); }
    ~Lock() {
      if (locked) o.mu.unlock();
    }
  };
};

inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(WFK_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return WFK_OK;
}

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern EncodeTiledFn g_encode_tiled;

}  // namespace wfk
