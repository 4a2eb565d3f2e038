// VIL frame staging: uint8 NHWT (SEVIR on-disk layout) -> normalised float N T (C=1) H W.
// Replaces SEVIRDataLoader.preprocess_data_dict + change_layout (reference
// pipeline/datasets/sevir/sevir.py:626-666, 88-101; cast at :587-592), bit-exactly:
//   out = fl32(scale) * ((float)u8 + fl32(offset))      rescale '01': scale = 1/255, offset = 0 (sevir.py:54-63);
//                                                        rescale 'sevir': scale = 1/47.54, offset = -33.44 (:44-53)
// HBM-bound: 128-bit loads of the byte stream into shared memory (the T=25 innermost bytes are
// not 16-byte aligned per pixel, so the transpose happens in smem), 128-bit stores per frame plane.
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

constexpr int kStagePix = 512;  // pixels per block
constexpr float kScale01 = 1.0f / 255.0f;  // == fl32(1/255) = 0x3b808081

// `windows` (optional): per output sequence (event index, first raw frame) into an event tensor [E, hw, t_src]
// (SEVIR events hold t_src = 49 frames; a sequence is the slice [t0, t0 + t), sevir.py:851-889). Without it the
// input is the already-sliced batch [n, hw, t] (t_src == t, t0 == 0).
template <bool HALF_OUT>
__global__ void __launch_bounds__(256) stage_vil_kernel(const uint8_t* __restrict__ in, int hw, int t, int t_src,
                                                        const int32_t* __restrict__ windows, float scale, float offset,
                                                        void* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t s_bytes[];  // [kStagePix * t_src]
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * kStagePix;
  const int npix = min(kStagePix, hw - p0);
  const int nbytes = npix * t_src;
  const int ev = windows ? windows[2 * n] : n;
  const int t0 = windows ? windows[2 * n + 1] : 0;
  const uint8_t* src = in + (static_cast<int64_t>(ev) * hw + p0) * t_src;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nvec = nbytes >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(s_bytes);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (int i = (nvec << 4) + threadIdx.x; i < nbytes; i += blockDim.x) s_bytes[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s_bytes[i] = src[i];
  }
  __syncthreads();
  // thread -> (frame index tq, group of 4 consecutive pixels pg)
  const int groups = (npix + 3) >> 2;
  for (int item = threadIdx.x; item < groups * t; item += blockDim.x) {
    const int ti = item / groups, pg = (item - ti * groups) << 2;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = pg + j;
      v[j] = (p < npix) ? __fmul_rn(__fadd_rn(static_cast<float>(s_bytes[p * t_src + t0 + ti]), offset), scale) : 0.f;
    }
    const int64_t o = (static_cast<int64_t>(n) * t + ti) * hw + p0 + pg;
    if (pg + 3 < npix && ((o & 3) == 0)) {
      if (HALF_OUT) {
        __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(static_cast<__half*>(out) + o) = u;
      } else {
        *reinterpret_cast<float4*>(static_cast<float*>(out) + o) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      for (int j = 0; j < 4 && pg + j < npix; ++j) {
        if (HALF_OUT) static_cast<__half*>(out)[o + j] = __float2half_rn(v[j]);
        else static_cast<float*>(out)[o + j] = v[j];
      }
    }
  }
}

// Fast path for whole sequences (no windowing) with a compile-time frame count T and hw % 4 == 0: a thread owns 4
// consecutive pixels = 4*T contiguous bytes = T aligned 32-bit words of the staged block, so every byte is a
// compile-time (word, lane) pair -- one LDS.32 per 4 outputs, one byte-lane I2F and one FMUL per output, one 128-bit
// (fp32) / 64-bit (fp16) store per frame and thread; consecutive threads store consecutive 16 bytes (512 B per warp).
// The generic kernel above spends ~6 instructions per output (a byte load each) and runs at 0.45-0.75 of the HBM
// roofline; this one is bound by the copy itself.
template <int T, bool HALF_OUT>
__global__ void __launch_bounds__(128) stage_vil_seq_kernel(const uint8_t* __restrict__ in, int hw, float scale, float offset,
                                                            void* __restrict__ out) {
  __shared__ __align__(16) uint32_t s_words[128 * T];  // 512 pixels x T bytes
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * kStagePix;
  const int npix = min(kStagePix, hw - p0);             // multiple of 4
  const int nwords = (npix * T) >> 2;
  const uint8_t* src = in + (static_cast<int64_t>(n) * hw + p0) * T;   // (hw * T) % 4 == 0 and p0 % 512 == 0: 16-byte aligned
  {
    const int nvec = nwords >> 2;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(s_words);
    for (int i = threadIdx.x; i < nvec; i += 128) d4[i] = __ldg(s4 + i);
    for (int i = (nvec << 2) + threadIdx.x; i < nwords; i += 128) s_words[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
  }
  __syncthreads();
  const int q = threadIdx.x;          // pixel group: pixels 4q .. 4q+3
  if (4 * q >= npix) return;
  uint32_t wd[T];
#pragma unroll
  for (int i = 0; i < T; ++i) wd[i] = s_words[q * T + i];   // word stride T (odd for T = 25): conflict-free
  const int64_t o0 = static_cast<int64_t>(n) * T * hw + p0 + 4 * q;
  const bool add = offset != 0.f;
#pragma unroll
  for (int ti = 0; ti < T; ++ti) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = j * T + ti;                             // byte index inside the thread's 4*T bytes
      float x = static_cast<float>((wd[b >> 2] >> (8 * (b & 3))) & 0xffu);
      if (add) x = __fadd_rn(x, offset);
      v[j] = __fmul_rn(x, scale);
    }
    const int64_t o = o0 + static_cast<int64_t>(ti) * hw;
    if (HALF_OUT) {
      __half2 a = __floats2half2_rn(v[0], v[1]), b2 = __floats2half2_rn(v[2], v[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b2);
      *reinterpret_cast<uint2*>(static_cast<__half*>(out) + o) = u;
    } else {
      *reinterpret_cast<float4*>(static_cast<float*>(out) + o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

}  // namespace wfk

extern "C" int wfk_stage_vil_u8(const uint8_t* nhwt, int n, int h, int w, int t, void* out_ntchw, int out_dtype,
                                void* stream) {
  WFK_ENTER(stream, nhwt);
  WFK_REQUIRE(nhwt && out_ntchw, "null pointer");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && t > 0 && t <= 64, "unsupported shape n=%d h=%d w=%d t=%d", n, h, w, t);
  WFK_REQUIRE(out_dtype == 0 || out_dtype == 1, "out_dtype must be 0 (f32) or 1 (f16)");
  const int hw = h * w;
  dim3 grid((hw + wfk::kStagePix - 1) / wfk::kStagePix, n);
  const size_t smem = static_cast<size_t>(wfk::kStagePix) * t;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uintptr_t align = reinterpret_cast<uintptr_t>(nhwt) | reinterpret_cast<uintptr_t>(out_ntchw);
  if (t == 25 && hw % 4 == 0 && (align & 15) == 0) {   // the Path-B sequence length (13 in + 12 out)
    if (out_dtype == 1) wfk::stage_vil_seq_kernel<25, true><<<grid, 128, 0, s>>>(nhwt, hw, wfk::kScale01, 0.f, out_ntchw);
    else wfk::stage_vil_seq_kernel<25, false><<<grid, 128, 0, s>>>(nhwt, hw, wfk::kScale01, 0.f, out_ntchw);
    return wfk::launched("stage_vil_seq_kernel");
  }
  if (out_dtype == 1) wfk::stage_vil_kernel<true><<<grid, 256, smem, s>>>(nhwt, hw, t, t, nullptr, wfk::kScale01, 0.f, out_ntchw);
  else wfk::stage_vil_kernel<false><<<grid, 256, smem, s>>>(nhwt, hw, t, t, nullptr, wfk::kScale01, 0.f, out_ntchw);
  return wfk::launched("stage_vil_kernel");
}

extern "C" int wfk_stage_vil_windows_ex(const uint8_t* events, int num_events, int h, int w, int t_raw, const int32_t* windows,
                                        int n, int t, float scale, float offset, void* out_ntchw, int out_dtype,
                                        void* stream) {
  WFK_ENTER(stream, events);
  WFK_REQUIRE(events && windows && out_ntchw, "null pointer");
  WFK_REQUIRE(num_events > 0 && n > 0 && n <= 65535 && h > 0 && w > 0, "unsupported shape");
  WFK_REQUIRE(t > 0 && t <= t_raw && t_raw <= 64, "need 0 < t (%d) <= t_raw (%d) <= 64", t, t_raw);
  WFK_REQUIRE(out_dtype == 0 || out_dtype == 1, "out_dtype must be 0 (f32) or 1 (f16)");
  const int hw = h * w;
  dim3 grid((hw + wfk::kStagePix - 1) / wfk::kStagePix, n);
  const size_t smem = static_cast<size_t>(wfk::kStagePix) * t_raw;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_dtype == 1) wfk::stage_vil_kernel<true><<<grid, 256, smem, s>>>(events, hw, t, t_raw, windows, scale, offset, out_ntchw);
  else wfk::stage_vil_kernel<false><<<grid, 256, smem, s>>>(events, hw, t, t_raw, windows, scale, offset, out_ntchw);
  return wfk::launched("stage_vil_kernel(windows)");
}

extern "C" int wfk_stage_vil_windows(const uint8_t* events, int num_events, int h, int w, int t_raw, const int32_t* windows,
                                     int n, int t, void* out_ntchw, int out_dtype, void* stream) {
  return wfk_stage_vil_windows_ex(events, num_events, h, w, t_raw, windows, n, t, wfk::kScale01, 0.f, out_ntchw, out_dtype,
                                  stream);
}
