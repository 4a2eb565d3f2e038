// Stem / head kernels of the two auxiliary conv families: layers whose channel
// counts are too thin for a 128 x N tensor-core tile and which are bandwidth-bound anyway.
//   Conv2d(1, C, 4, stride 2, padding 1) (+ folded BatchNorm) + activation
//       PosAwareAE_TF.enc[0].down (pipeline/models/ae_64x8x8_lin.py:31-34), NLayerDiscriminator.main[0..1]
//       (pipeline/models/autoencoderkl/losses/model.py:121)
//   Conv2d(C, 1, 1, padding 1)            NLayerDiscriminator logit head (losses/model.py:149-150)
//   hinge / mean reductions of the logits (losses/contperceptual.py:19-23)
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

__device__ __forceinline__ float aux_act(int kind, float x, float slope) {
  switch (kind) {
    case WFK_ACT_LEAKY_RELU: return x > 0.f ? x : slope * x;
    case WFK_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case WFK_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case WFK_ACT_SILU: return x / (1.f + __expf(-x));
    default: return x;
  }
}

constexpr int kStemThreads = 256;
constexpr int kStemPixPerBlock = 512;

// One thread = 8 output channels, looping over output pixels; its 16 x 8 weights live in REGISTERS (the first version
// re-read them from shared memory for every pixel: 4 x LDS.128 per output value, shared-memory-bandwidth bound at 10x
// the layer's HBM time -- 36 % of the discriminator's forward pass), the 16 inputs of a pixel are read once per thread
// (L1 broadcast among the 8 threads sharing the pixel). 128 FMAs per 16 input loads and one 128-bit store.
__global__ void __launch_bounds__(kStemThreads) conv4x4s2_c1in_kernel(
    const float* __restrict__ in, int h, int w, const float* __restrict__ wt, const float* __restrict__ bias, int cout,
    int act, float slope, __half* __restrict__ out, int act2, const float* __restrict__ scale2,
    const float* __restrict__ shift2, __half* __restrict__ out2) {
  const int n = blockIdx.y;
  const int oh = h >> 1, ow = w >> 1;
  const int octets = cout >> 3;
  const int oc = (threadIdx.x % octets) << 3;
  const int pl = threadIdx.x / octets;
  const int pstep = kStemThreads / octets;
  float wr[16][8], br[8], s2[8], h2v[8];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + t * cout + oc));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + t * cout + oc + 4));
    wr[t][0] = w0.x; wr[t][1] = w0.y; wr[t][2] = w0.z; wr[t][3] = w0.w;
    wr[t][4] = w1.x; wr[t][5] = w1.y; wr[t][6] = w1.z; wr[t][7] = w1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = __ldg(bias + oc + j);
    s2[j] = (out2 != nullptr && scale2) ? __ldg(scale2 + oc + j) : 1.f;
    h2v[j] = (out2 != nullptr && shift2) ? __ldg(shift2 + oc + j) : 0.f;
  }
  const float* inn = in + static_cast<int64_t>(n) * h * w;
  const int p_begin = blockIdx.x * kStemPixPerBlock;
  const int p_end = min(oh * ow, p_begin + kStemPixPerBlock);
  // (y, x) of this thread's pixel are advanced incrementally and interior windows are read through one row pointer with
  // immediate offsets: the first version spent ~480 integer instructions per pixel on a division, 16 bounds checks and
  // 16 64-bit address computations next to its 128 FMAs (ncu: FFMA 21 % of 1.08 G warp instructions)
  int p = p_begin + pl;
  int y = p / ow, x = p - y * ow;
#pragma unroll 1
  for (; p < p_end; p += pstep) {
    float z[16];
    if (y > 0 && y < oh - 1 && x > 0 && x < ow - 1) {
      const float* r0 = inn + (2 * y - 1) * w + (2 * x - 1);
#pragma unroll
      for (int t = 0; t < 16; ++t) z[t] = __ldg(r0 + (t >> 2) * w + (t & 3));
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int yy = 2 * y - 1 + (t >> 2), xx = 2 * x - 1 + (t & 3);
        z[t] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(inn + yy * w + xx) : 0.f;
      }
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wr[t][j], z[t], acc[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = aux_act(act, acc[j], slope);
    const int64_t o = (static_cast<int64_t>(n) * oh * ow + p) * cout + oc;
    uint4 u;
    __half2* h2 = reinterpret_cast<__half2*>(&u);
    if (out != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
      *reinterpret_cast<uint4*>(out + o) = u;
    }
    if (out2 != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h2[e] = __floats2half2_rn(aux_act(act2, fmaf(acc[2 * e], s2[2 * e], h2v[2 * e]), slope),
                                  aux_act(act2, fmaf(acc[2 * e + 1], s2[2 * e + 1], h2v[2 * e + 1]), slope));
      *reinterpret_cast<uint4*>(out2 + o) = u;
    }
    x += pstep;
    while (x >= ow) {
      x -= ow;
      ++y;
    }
  }
}

// One warp = one output pixel of the padded map: dot product over cin (fp16 NHWC row), or the bare bias.
__global__ void __launch_bounds__(256) conv1x1_cout1_kernel(const __half* __restrict__ in, int n, int h, int w, int cin,
                                                           const float* __restrict__ wt, float bias, int pad,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int oh = h + 2 * pad, ow = w + 2 * pad;
  const int64_t total = static_cast<int64_t>(n) * oh * ow;
  for (int64_t o = warp_global; o < total; o += nwarps) {
    const int fn = static_cast<int>(o / (oh * ow));
    const int r = static_cast<int>(o - static_cast<int64_t>(fn) * oh * ow);
    const int y = r / ow - pad, x = r % ow - pad;
    float acc = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) {
      const __half* row = in + ((static_cast<int64_t>(fn) * h + y) * w + x) * cin;
      for (int c = lane * 8; c < cin; c += 256) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + c));
        const __half2* a2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(a2[e]);
          acc = fmaf(f.x, __ldg(wt + c + 2 * e), acc);
          acc = fmaf(f.y, __ldg(wt + c + 2 * e + 1), acc);
        }
      }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o] = acc + bias;
  }
}

__global__ void __launch_bounds__(256) logit_sums_kernel(const float* __restrict__ x, int64_t count,
                                                        double* __restrict__ sums) {
  __shared__ double s_red[3][8];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    a += v;
    b += fmaxf(1.f - v, 0.f);
    c += fmaxf(1.f + v, 0.f);
  }
  for (int s = 16; s; s >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, s);
    b += __shfl_xor_sync(0xffffffffu, b, s);
    c += __shfl_xor_sync(0xffffffffu, c, s);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = a;
    s_red[1][threadIdx.x >> 5] = b;
    s_red[2][threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_red[threadIdx.x][i];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

}  // namespace wfk

extern "C" int wfk_conv4x4s2_c1in(const float* in, int n, int h, int w, const float* weight, const float* bias,
                                  int cout, int act, float act_slope, void* out, int act2, const float* scale2,
                                  const float* shift2, void* out2, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && weight && bias && (out || out2), "null pointer");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "bad shape %dx%dx%d", n, h, w);
  WFK_REQUIRE(cout % 8 == 0 && cout >= 8 && cout <= 2048 && wfk::kStemThreads % (cout / 8) == 0, "cout=%d unsupported", cout);
  const size_t smem = 0;   // weights, bias and the second affine map live in registers
  WFK_REQUIRE((reinterpret_cast<uintptr_t>(weight) & 15) == 0, "weight must be 16-byte aligned");
  dim3 grid(((h / 2) * (w / 2) + wfk::kStemPixPerBlock - 1) / wfk::kStemPixPerBlock, n);
  wfk::conv4x4s2_c1in_kernel<<<grid, wfk::kStemThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      in, h, w, weight, bias, cout, act, act_slope, static_cast<__half*>(out), act2, scale2, shift2,
      static_cast<__half*>(out2));
  return wfk::launched("conv4x4s2_c1in_kernel");
}

extern "C" int wfk_conv1x1_cout1(const void* in, int n, int h, int w, int cin, const float* weight, float bias, int pad,
                                 float* out, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && weight && out, "null pointer");
  WFK_REQUIRE(n > 0 && h > 0 && w > 0 && cin > 0 && cin % 8 == 0 && pad >= 0, "bad shape");
  const int64_t total = static_cast<int64_t>(n) * (h + 2 * pad) * (w + 2 * pad);
  int64_t blocks = (total + 7) / 8;
  const int64_t cap = static_cast<int64_t>(wfk::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  wfk::conv1x1_cout1_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in), n, h, w, cin, weight, bias, pad, out);
  return wfk::launched("conv1x1_cout1_kernel");
}

extern "C" int wfk_logit_sums(const float* x, int64_t count, double* sums, void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && sums && count > 0, "bad argument");
  int64_t blocks = (count + 255) / 256;
  const int64_t cap = static_cast<int64_t>(wfk::num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  wfk::logit_sums_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, count, sums);
  return wfk::launched("logit_sums_kernel");
}
