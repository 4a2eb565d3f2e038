// decoder tail: GroupNorm + SiLU + conv3x3(C -> 1) fused in one pass over the raw fp16 stream.
// Replaces conv_norm_out / conv_act / conv_out of Decoder.forward (reference
// pipeline/models/autoencoderkl/vae.py:162-164) -- previously a GroupNorm pass (read + write of the
// 128-channel 384x384 tensor) followed by a gather-style direct convolution.
//
// out[p] = bias + sum_tap sum_c w[tap][c] * silu(a[c]*x[p+tap][c] + b[c])   (zero outside the image)
// is evaluated as d[q][tap] = sum_c w[tap][c] * act(x[q][c]) for every pixel q of a 16x16 tile plus a
// 1-pixel halo (each input pixel is read ONCE, 256 contiguous bytes, by one thread), followed by the
// 9-value gather out[p] = sum_tap d[p + tap][tap] through shared memory. HBM-bound: C*2 bytes per pixel.
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

constexpr int kOT = 16;                 // output tile edge
constexpr int kOH = kOT + 2;            // with halo
constexpr int kOutThreads = 352;        // >= kOH*kOH = 324

__global__ void __launch_bounds__(kOutThreads) gn_silu_conv3x3_c1_kernel(
    const __half* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, int h, int w, int c, int groups, float eps, const float* __restrict__ wt,
    float bias, float* __restrict__ out) {
  extern __shared__ float s_mem[];
  float* s_a = s_mem;                // [c]
  float* s_b = s_a + c;              // [c]
  float* s_w = s_b + c;              // [9][c]
  float* s_d = s_w + 9 * c;          // [kOH*kOH][9]
  float* s_mean = s_d + kOH * kOH * 9;
  float* s_rstd = s_mean + groups;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kOT, y0 = blockIdx.y * kOT;
  const int cpg = c / groups;
  if (threadIdx.x < groups) {
    const double cnt = static_cast<double>(cpg) * h * w;
    const int g = threadIdx.x;
    const double sum = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0];
    const double sq = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
  }
  for (int i = threadIdx.x; i < 9 * c; i += blockDim.x) s_w[i] = wt[i];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float a = s_rstd[ch / cpg] * gamma[ch];
    s_a[ch] = a;
    s_b[ch] = beta[ch] - s_mean[ch / cpg] * a;
  }
  __syncthreads();
  if (threadIdx.x < kOH * kOH) {
    const int qy = threadIdx.x / kOH, qx = threadIdx.x - qy * kOH;
    const int y = y0 + qy - 1, xx = x0 + qx - 1;
    float d[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) d[t] = 0.f;
    if (y >= 0 && y < h && xx >= 0 && xx < w) {
      const uint4* px = reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * h + y) * w + xx) * c);
      for (int v = 0; v < (c >> 3); ++v) {
        const uint4 u = __ldg(px + v);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        float act[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          const float z0 = fmaf(f.x, s_a[8 * v + 2 * e], s_b[8 * v + 2 * e]);
          const float z1 = fmaf(f.y, s_a[8 * v + 2 * e + 1], s_b[8 * v + 2 * e + 1]);
          act[2 * e] = __fdividef(z0, 1.f + __expf(-z0));
          act[2 * e + 1] = __fdividef(z1, 1.f + __expf(-z1));
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * c + 8 * v);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + t * c + 8 * v + 4);
          d[t] = fmaf(act[0], w0.x, d[t]);
          d[t] = fmaf(act[1], w0.y, d[t]);
          d[t] = fmaf(act[2], w0.z, d[t]);
          d[t] = fmaf(act[3], w0.w, d[t]);
          d[t] = fmaf(act[4], w1.x, d[t]);
          d[t] = fmaf(act[5], w1.y, d[t]);
          d[t] = fmaf(act[6], w1.z, d[t]);
          d[t] = fmaf(act[7], w1.w, d[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_d[threadIdx.x * 9 + t] = d[t];
  }
  __syncthreads();
  if (threadIdx.x < kOT * kOT) {
    const int py = threadIdx.x / kOT, pxx = threadIdx.x - py * kOT;
    const int y = y0 + py, xx = x0 + pxx;
    if (y < h && xx < w) {
      float acc = bias;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) acc += s_d[((py + r) * kOH + (pxx + s)) * 9 + r * 3 + s];
      out[(static_cast<int64_t>(n) * h + y) * w + xx] = acc;
    }
  }
}

}  // namespace wfk

extern "C" int wfk_gn_silu_conv3x3_c1(const void* x, const double* stats, const float* gamma, const float* beta, int n,
                                      int h, int w, int c, int groups, float eps, const float* weight, float bias,
                                      float* out, void* stream) {
  WFK_REQUIRE_INIT();
  WFK_REQUIRE(x && stats && gamma && beta && weight && out, "null pointer");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0, "bad shape");
  WFK_REQUIRE(c % 8 == 0 && c > 0 && c <= 1024 && c % groups == 0 && groups <= 256, "unsupported c=%d groups=%d", c, groups);
  const size_t smem = (static_cast<size_t>(c) * 11 + wfk::kOH * wfk::kOH * 9 + 2 * groups) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::gn_silu_conv3x3_c1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  WFK_REQUIRE(smem <= 100 * 1024, "channel count too large for shared memory");
  dim3 grid((w + wfk::kOT - 1) / wfk::kOT, (h + wfk::kOT - 1) / wfk::kOT, n);
  wfk::gn_silu_conv3x3_c1_kernel<<<grid, wfk::kOutThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), stats, gamma, beta, h, w, c, groups, eps, weight, bias, out);
  return wfk::launched("gn_silu_conv3x3_c1_kernel");
}
