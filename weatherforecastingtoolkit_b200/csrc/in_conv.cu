// encoder.conv_in: 3x3 (pad 1) convolution 1 -> C channels of the staged fp32 frame, writing the NHWC fp16
// stream and its GroupNorm statistics (reference pipeline/models/autoencoderkl/vae.py:24, 72).
// Write-bound (2*C bytes per pixel): each thread keeps its 8 output channels' 9x8 weights in registers and
// walks a strip of pixels; the 16 threads that share a pixel read the same 9 inputs (L1 broadcast).
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

constexpr int kInThreads = 256;
constexpr int kInPixPerBlock = 1024;

__global__ void __launch_bounds__(kInThreads) conv3x3_c1in_kernel(const float* __restrict__ in, int h, int w,
                                                                  const float* __restrict__ wt,  // [9][cout]
                                                                  const float* __restrict__ bias, int cout,
                                                                  __half* __restrict__ out, double* __restrict__ stats,
                                                                  int cpg) {
  __shared__ float s_part[kInThreads][4];
  const int n = blockIdx.y;
  const int octets = cout >> 3;
  const int oc = (threadIdx.x % octets) << 3;
  const int pl = threadIdx.x / octets;
  const int pstep = kInThreads / octets;
  const int hw = h * w;
  float wreg[9][8], breg[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wt + t * cout + oc));
    const float4 b = __ldg(reinterpret_cast<const float4*>(wt + t * cout + oc + 4));
    wreg[t][0] = a.x; wreg[t][1] = a.y; wreg[t][2] = a.z; wreg[t][3] = a.w;
    wreg[t][4] = b.x; wreg[t][5] = b.y; wreg[t][6] = b.z; wreg[t][7] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) breg[j] = bias[oc + j];
  float ps[2] = {0.f, 0.f}, pq[2] = {0.f, 0.f};
  const float* inn = in + static_cast<int64_t>(n) * hw;
  const int p_begin = blockIdx.x * kInPixPerBlock;
  const int p_end = min(hw, p_begin + kInPixPerBlock);
  for (int p = p_begin + pl; p < p_end; p += pstep) {
    const int y = p / w, x = p - y * w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = breg[j];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const float z = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(inn + yy * w + xx) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wreg[t][j], z, acc[j]);
    }
    uint4 u;
    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
    *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(n) * hw + p) * cout + oc) = u;
    if (cpg >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ps[0] += acc[j];
        pq[0] = fmaf(acc[j], acc[j], pq[0]);
      }
    } else {
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ps[g] += acc[4 * g + j];
          pq[g] = fmaf(acc[4 * g + j], acc[4 * g + j], pq[g]);
        }
    }
  }
  if (stats != nullptr) {
    const int groups = cout / cpg;
    s_part[threadIdx.x][0] = ps[0];
    s_part[threadIdx.x][1] = pq[0];
    s_part[threadIdx.x][2] = ps[1];
    s_part[threadIdx.x][3] = pq[1];
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {  // fixed-order fold (deterministic)
      double s = 0.0, q = 0.0;
      if (cpg >= 8) {
        const int o_begin = (g * cpg) >> 3, o_end = ((g + 1) * cpg) >> 3;
        for (int t = 0; t < kInThreads; ++t) {
          const int o = t % octets;
          if (o >= o_begin && o < o_end) {
            s += s_part[t][0];
            q += s_part[t][1];
          }
        }
      } else {
        const int o = g >> 1, sub = g & 1;
        for (int t = o; t < kInThreads; t += octets) {
          s += s_part[t][2 * sub];
          q += s_part[t][2 * sub + 1];
        }
      }
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0], s);
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1], q);
    }
  }
}

}  // namespace wfk

// Called by wfk_conv3x3_small_cin for cin == 1 without a pre-conv.
int wfk_launch_c1in(const float* in, int n, int h, int w, const float* weight, const float* bias, int cout, void* out,
                    double* stats, int cpg, cudaStream_t s) {
  dim3 grid((h * w + wfk::kInPixPerBlock - 1) / wfk::kInPixPerBlock, n);
  wfk::conv3x3_c1in_kernel<<<grid, wfk::kInThreads, 0, s>>>(in, h, w, weight, bias, cout, static_cast<__half*>(out),
                                                           stats, cpg);
  return wfk::launched("conv3x3_c1in_kernel");
}
