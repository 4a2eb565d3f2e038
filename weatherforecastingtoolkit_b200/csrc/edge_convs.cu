// Direct (CUDA-core) 3x3 convolutions for the four layers whose channel counts are too thin for a
// 128 x N tensor-core tile and which are bandwidth-bound anyway (SURVEY section 7 "edge convs"):
//   encoder.conv_in   1 -> 128   (vae.py:24)          small_cin
//   decoder.conv_in   4 -> 512   (vae.py:103) with post_quant_conv 1x1 folded in front (autoencoder_kl.py:87)
//   encoder.conv_out  512 -> 8   (vae.py:68)  with quant_conv 1x1 folded behind (autoencoder_kl.py:82)
//   decoder.conv_out  128 -> 1   (vae.py:148)         small_cout
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

constexpr int kMaxCin = 4;

// in: [n, cin, h, w] fp32; wt: [cin*9][cout] fp32 (tap-major, cout contiguous); out: NHWC fp16.
// One thread = one pixel x 8 output channels. Also accumulates GroupNorm (sum, sumsq) per
// (frame, group) of the fp32 result.
__global__ void __launch_bounds__(256) conv3x3_small_cin_kernel(
    const float* __restrict__ in, int cin, int h, int w, const float* __restrict__ pre_w,
    const float* __restrict__ pre_b, const float* __restrict__ wt, const float* __restrict__ bias, int cout,
    __half* __restrict__ out, double* __restrict__ stats, int cpg, int pix_per_block) {
  __shared__ float s_part[256][4];  // per-thread (sum, sumsq) x up to 2 groups: reduced in a fixed order
  const int n = blockIdx.y;
  const int octets = cout >> 3;
  const int groups = stats ? cout / cpg : 0;
  float ps[2] = {0.f, 0.f}, pq[2] = {0.f, 0.f};
  const int hw = h * w;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  const int oc = (threadIdx.x % octets) << 3;
  const int pl = threadIdx.x / octets;
  const int pstep = blockDim.x / octets;
  const float* inn = in + static_cast<int64_t>(n) * cin * hw;
  for (int p = p_begin + pl; p < p_end; p += pstep) {
    const int y = p / w, x = p - y * w;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias[oc + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
      float z[kMaxCin];
#pragma unroll
      for (int ci = 0; ci < kMaxCin; ++ci)
        z[ci] = (ci < cin) ? __ldg(inn + static_cast<int64_t>(ci) * hw + yy * w + xx) : 0.f;
      if (pre_w != nullptr) {
        float z2[kMaxCin];
#pragma unroll
        for (int co = 0; co < kMaxCin; ++co) {
          float a = 0.f;
          if (co < cin) {
            a = pre_b[co];
#pragma unroll
            for (int ci = 0; ci < kMaxCin; ++ci)
              if (ci < cin) a = fmaf(pre_w[co * cin + ci], z[ci], a);
          }
          z2[co] = a;
        }
#pragma unroll
        for (int ci = 0; ci < kMaxCin; ++ci) z[ci] = z2[ci];
      }
#pragma unroll
      for (int ci = 0; ci < kMaxCin; ++ci) {
        if (ci < cin) {
          const float4* wp = reinterpret_cast<const float4*>(wt + static_cast<int64_t>(ci * 9 + tap) * cout + oc);
          const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
          acc[0] = fmaf(w0.x, z[ci], acc[0]);
          acc[1] = fmaf(w0.y, z[ci], acc[1]);
          acc[2] = fmaf(w0.z, z[ci], acc[2]);
          acc[3] = fmaf(w0.w, z[ci], acc[3]);
          acc[4] = fmaf(w1.x, z[ci], acc[4]);
          acc[5] = fmaf(w1.y, z[ci], acc[5]);
          acc[6] = fmaf(w1.z, z[ci], acc[6]);
          acc[7] = fmaf(w1.w, z[ci], acc[7]);
        }
      }
    }
    uint4 u;
    __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
    *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(n) * hw + p) * cout + oc) = u;
    if (stats != nullptr) {
      if (cpg >= 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ps[0] += acc[j];
          pq[0] = fmaf(acc[j], acc[j], pq[0]);
        }
      } else {  // cpg == 4: two groups per octet
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ps[g] += acc[4 * g + j];
            pq[g] = fmaf(acc[4 * g + j], acc[4 * g + j], pq[g]);
          }
        }
      }
    }
  }
  if (stats != nullptr) {
    s_part[threadIdx.x][0] = ps[0];
    s_part[threadIdx.x][1] = pq[0];
    s_part[threadIdx.x][2] = ps[1];
    s_part[threadIdx.x][3] = pq[1];
    __syncthreads();
    // thread g folds every contributor of group g in thread-index order (deterministic)
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      double s = 0.0, q = 0.0;
      if (cpg >= 8) {
        const int o_begin = (g * cpg) >> 3, o_end = ((g + 1) * cpg) >> 3;
        for (int t = 0; t < static_cast<int>(blockDim.x); ++t) {
          const int o = t % octets;
          if (o >= o_begin && o < o_end) {
            s += s_part[t][0];
            q += s_part[t][1];
          }
        }
      } else {
        const int o = g >> 1, sub = g & 1;
        for (int t = o; t < static_cast<int>(blockDim.x); t += octets) {
          s += s_part[t][2 * sub];
          q += s_part[t][2 * sub + 1];
        }
      }
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0], s);
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1], q);
    }
  }
}

// in: [n, h, w, cin] fp16 (already GroupNorm+SiLU'd); wt: [COUT][9][cin] fp16; out: [n, COUT, h, w] fp32.
// One warp = one output pixel; lanes split the 9 x cin/8 (tap, 8-channel vector) items.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_small_cout_kernel(const __half* __restrict__ in, int n, int h, int w,
                                                                 int cin, const __half* __restrict__ wt,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ post_w,
                                                                 const float* __restrict__ post_b,
                                                                 float* __restrict__ out, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t hw = static_cast<int64_t>(h) * w;
  const int64_t total = hw * n;
  const int vpp = cin >> 3;
  const int items = 9 * vpp;
  for (int64_t pix = warp_global; pix < total; pix += nwarps) {
    const int fn = static_cast<int>(pix / hw);
    const int p = static_cast<int>(pix - fn * hw);
    const int y = p / w, x = p - y * w;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
    for (int it = lane; it < items; it += 32) {
      const int tap = it / vpp, cv = (it - tap * vpp) << 3;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<int64_t>(fn) * h + yy) * w + xx) * cin + cv));
      const __half2* a2 = reinterpret_cast<const __half2*>(&u);
      float a[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(a2[e]);
        a[2 * e] = f.x;
        a[2 * e + 1] = f.y;
      }
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        const uint4 wu = __ldg(reinterpret_cast<const uint4*>(wt + (static_cast<int64_t>(co) * 9 + tap) * cin + cv));
        const __half2* w2 = reinterpret_cast<const __half2*>(&wu);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(w2[e]);
          acc[co] = fmaf(a[2 * e], f.x, acc[co]);
          acc[co] = fmaf(a[2 * e + 1], f.y, acc[co]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
      acc[co] += bias[co];
    }
    if (lane < COUT) {
      float r;
      if (post_w != nullptr) {
        r = post_b[lane];
#pragma unroll
        for (int co = 0; co < COUT; ++co) r = fmaf(post_w[lane * COUT + co], acc[co], r);
      } else {
        r = 0.f;
#pragma unroll
        for (int co = 0; co < COUT; ++co) r = (co == lane) ? acc[co] : r;
      }
      if (act == WFK_ACT_SIGMOID) r = 1.f / (1.f + __expf(-r));  // nn.Sigmoid tail (ae_64x8x8_lin.py:85,105)
      out[(static_cast<int64_t>(fn) * COUT + lane) * hw + p] = r;
    }
  }
}

}  // namespace wfk

int wfk_launch_c1in(const float* in, int n, int h, int w, const float* weight, const float* bias, int cout, void* out,
                    double* stats, int cpg, cudaStream_t s);

extern "C" int wfk_conv3x3_small_cin(const float* in, int n, int cin, int h, int w, const float* pre_w,
                                     const float* pre_b, const float* weight, const float* bias, int cout, void* out,
                                     double* stats, int cpg, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && weight && bias && out, "null pointer");
  WFK_REQUIRE(cin >= 1 && cin <= wfk::kMaxCin, "cin=%d unsupported (1..%d)", cin, wfk::kMaxCin);
  WFK_REQUIRE(cout % 8 == 0 && cout >= 8 && cout <= 2048 && 256 % (cout / 8) == 0, "cout=%d unsupported", cout);
  WFK_REQUIRE((pre_w == nullptr) == (pre_b == nullptr), "pre_w / pre_b must both be given or both NULL");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0, "bad shape");
  if (stats) WFK_REQUIRE(cpg >= 4 && cout % cpg == 0 && (cpg == 4 || cpg % 8 == 0), "cpg=%d unsupported", cpg);
  if (cin == 1 && pre_w == nullptr && 256 % (cout / 8) == 0 && (!stats || cpg == 4 || cpg % 8 == 0))
    return wfk_launch_c1in(in, n, h, w, weight, bias, cout, out, stats, stats ? cpg : 8, static_cast<cudaStream_t>(stream));
  const int ppb = 64;
  dim3 grid((h * w + ppb - 1) / ppb, n);
  const size_t smem = 0;
  wfk::conv3x3_small_cin_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      in, cin, h, w, pre_w, pre_b, weight, bias, cout, static_cast<__half*>(out), stats, cpg, ppb);
  return wfk::launched("conv3x3_small_cin_kernel");
}

extern "C" int wfk_conv3x3_small_cout(const void* in, int n, int h, int w, int cin, const void* weight_h,
                                      const float* bias, int cout, const float* post_w, const float* post_b,
                                      float* out, void* stream) {
  return wfk_conv3x3_small_cout_act(in, n, h, w, cin, weight_h, bias, cout, post_w, post_b, WFK_ACT_NONE, out, stream);
}

extern "C" int wfk_conv3x3_small_cout_act(const void* in, int n, int h, int w, int cin, const void* weight_h,
                                          const float* bias, int cout, const float* post_w, const float* post_b,
                                          int act, float* out, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(act == WFK_ACT_NONE || act == WFK_ACT_SIGMOID, "act must be none or sigmoid");
  WFK_REQUIRE(in && weight_h && bias && out, "null pointer");
  WFK_REQUIRE(cin % 8 == 0 && cin > 0, "cin must be a multiple of 8");
  WFK_REQUIRE((post_w == nullptr) == (post_b == nullptr), "post_w / post_b must both be given or both NULL");
  WFK_REQUIRE(n > 0 && h > 0 && w > 0, "bad shape");
  const int64_t total = static_cast<int64_t>(n) * h * w;
  int64_t blocks = (total + 7) / 8;
  const int64_t cap = static_cast<int64_t>(wfk::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* ih = static_cast<const __half*>(in);
  const __half* wh = static_cast<const __half*>(weight_h);
#define WFK_LAUNCH_SC(CO)                                                                                       \
  wfk::conv3x3_small_cout_kernel<CO><<<static_cast<unsigned>(blocks), 256, 0, s>>>(ih, n, h, w, cin, wh, bias, \
                                                                                  post_w, post_b, out, act)
  switch (cout) {
    case 1: WFK_LAUNCH_SC(1); break;
    case 2: WFK_LAUNCH_SC(2); break;
    case 4: WFK_LAUNCH_SC(4); break;
    case 8: WFK_LAUNCH_SC(8); break;
    default: return wfk::fail(WFK_ERR_INVALID, "cout=%d unsupported (1, 2, 4, 8)", cout);
  }
#undef WFK_LAUNCH_SC
  return wfk::launched("conv3x3_small_cout_kernel");
}
