// Internal helpers shared by the translation units of libwfk_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/wfk_b200.h"

namespace wfk {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;
extern int g_device;           // -1 until wfk_init
extern int g_num_sms;

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

#define WFK_REQUIRE_INIT()                                                                  \
  do {                                                                                      \
    if (::wfk::g_device < 0) return ::wfk::fail(WFK_ERR_NOT_INIT, "wfk_init was not called"); \
  } while (0)

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
