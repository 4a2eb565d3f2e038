// ConvModel latent compressor: the whole per-frame network of
// experiments/v1_experiments/pretrained_ae_convae_sevir/train.py:58-143 in ONE kernel, one CTA per latent frame,
// every activation resident in shared memory (the largest is 8 x 48 x 48 fp32 = 72 KB; two ping-pong buffers):
//   ConvEncoder: conv3x3 4->8, LayerNorm([8,48,48]), LeakyReLU; 3 x [conv4x4 s2 p1 8->8, LayerNorm, LeakyReLU] (48->6)
//   to_latent Linear(288 -> 512)  -> z;  to_reconstruction Linear(512 -> 288)
//   ConvDecoder: 3 x [ConvTranspose4x4 s2 p1 8->8, LayerNorm, LeakyReLU] (6->48); conv3x3 8->4
// plus the HuberLoss(pred, input) partial sums of the experiment's validation_step (train.py:160, 193-194).
// The reference runs ~30 tiny library kernels per call (each far below one wave of a B200); here the model is a
// single launch and the only HBM traffic is the 36 KB frame in, 36 KB + 2 KB out and the (L2-resident) weights.
// All arithmetic is fp32 on the CUDA cores: the model is ~10 MFLOP per frame, latency- not throughput-bound.
#include "internal.h"

namespace wfk {

constexpr int kCmThreads = 512;
constexpr int kCmC = 8;        // bottleneck channels
constexpr float kCmSlope = 0.01f;  // nn.LeakyReLU() default

struct ConvModelWeights {
  const float* conv0_w;  // [8, cin, 3, 3]
  const float* conv0_b;  // [8]
  const float* ln_w[7];  // LayerNorm weight / bias of conv0, down1..3, up1..3: [8, H, W]
  const float* ln_b[7];
  const float* down_w[3];  // [8, 8, 4, 4]
  const float* down_b[3];
  const float* to_latent_w;  // [latent, 288]
  const float* to_latent_b;
  const float* to_rec_w;  // [288, latent]
  const float* to_rec_b;
  const float* up_w[3];  // ConvTranspose2d: [8 (in), 8 (out), 4, 4]
  const float* up_b[3];
  const float* out_w;  // [cout, 8, 3, 3]
  const float* out_b;
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kCmThreads / 32; ++i) t += s_red[i];  // same order in every thread: deterministic
  return t;
}

// nn.LayerNorm(normalized_shape=[C, H, W], eps=1e-5) with elementwise affine, then LeakyReLU, in place.
__device__ void layernorm_lrelu(float* x, int n, const float* __restrict__ w, const float* __restrict__ b, float* s_red) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kCmThreads) s += x[i];
  const float mean = block_sum(s, s_red) / static_cast<float>(n);
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += kCmThreads) {
    const float d = x[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(block_sum(q, s_red) / static_cast<float>(n) + 1e-5f);
  for (int i = threadIdx.x; i < n; i += kCmThreads) {
    const float y = (x[i] - mean) * rstd * __ldg(w + i) + __ldg(b + i);
    x[i] = y > 0.f ? y : kCmSlope * y;
  }
  __syncthreads();
}

// 3x3 stride 1 pad 1: in [ci, h, w] (shared or global) -> out [co, h, w]
__device__ void conv3x3(const float* in, int ci, int h, int w, const float* __restrict__ wt, const float* __restrict__ bias,
                        int co, float* out) {
  const int hw = h * w;
  for (int o = threadIdx.x; o < co * hw; o += kCmThreads) {
    const int c = o / hw, p = o - c * hw, y = p / w, x = p - y * w;
    float acc = __ldg(bias + c);
    for (int k = 0; k < ci; ++k)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yy = y + r - 1;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int xx = x + s - 1;
          if (xx < 0 || xx >= w) continue;
          acc = fmaf(in[k * hw + yy * w + xx], __ldg(wt + ((c * ci + k) * 3 + r) * 3 + s), acc);
        }
      }
    out[o] = acc;
  }
  __syncthreads();
}

// Conv2d(8, 8, 4, stride 2, padding 1): in [8, h, w] -> out [8, h/2, w/2]
__device__ void conv4x4s2(const float* in, int h, int w, const float* __restrict__ wt, const float* __restrict__ bias,
                          float* out) {
  const int oh = h >> 1, ow = w >> 1, ohw = oh * ow, hw = h * w;
  for (int o = threadIdx.x; o < kCmC * ohw; o += kCmThreads) {
    const int c = o / ohw, p = o - c * ohw, y = p / ow, x = p - y * ow;
    float acc = __ldg(bias + c);
    for (int k = 0; k < kCmC; ++k)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int yy = 2 * y - 1 + r;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int xx = 2 * x - 1 + s;
          if (xx < 0 || xx >= w) continue;
          acc = fmaf(in[k * hw + yy * w + xx], __ldg(wt + ((c * kCmC + k) * 4 + r) * 4 + s), acc);
        }
      }
    out[o] = acc;
  }
  __syncthreads();
}

// ConvTranspose2d(8, 8, 4, stride 2, padding 1): in [8, h, w] -> out [8, 2h, 2w]; weight [in, out, 4, 4].
// Output row Y receives input rows y with Y = 2y - 1 + r: r has the parity of Y + 1.
__device__ void convT4x4s2(const float* in, int h, int w, const float* __restrict__ wt, const float* __restrict__ bias,
                           float* out) {
  const int oh = 2 * h, ow = 2 * w, ohw = oh * ow, hw = h * w;
  for (int o = threadIdx.x; o < kCmC * ohw; o += kCmThreads) {
    const int c = o / ohw, p = o - c * ohw, Y = p / ow, X = p - Y * ow;
    float acc = __ldg(bias + c);
    for (int k = 0; k < kCmC; ++k)
#pragma unroll
      for (int ri = 0; ri < 2; ++ri) {
        const int r = ((Y + 1) & 1) + 2 * ri;
        const int y2 = Y + 1 - r;  // = 2y
        if (y2 < 0 || y2 >= 2 * h) continue;
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          const int s = ((X + 1) & 1) + 2 * si;
          const int x2 = X + 1 - s;
          if (x2 < 0 || x2 >= 2 * w) continue;
          acc = fmaf(in[k * hw + (y2 >> 1) * w + (x2 >> 1)], __ldg(wt + ((k * kCmC + c) * 4 + r) * 4 + s), acc);
        }
      }
    out[o] = acc;
  }
  __syncthreads();
}

// y[n_out] = W[n_out, n_in] x + b: one warp per output row (coalesced weight reads), x in shared memory
__device__ void linear(const float* x, int n_in, const float* __restrict__ wt, const float* __restrict__ bias, int n_out,
                       float* y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < n_out; o += kCmThreads / 32) {
    float acc = 0.f;
    for (int k = lane; k < n_in; k += 32) acc = fmaf(x[k], __ldg(wt + static_cast<int64_t>(o) * n_in + k), acc);
    for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) y[o] = acc + __ldg(bias + o);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCmThreads, 1) convmodel_kernel(const float* __restrict__ x, int cin, int latent_dim,
                                                                 ConvModelWeights W, float* __restrict__ z,
                                                                 float* __restrict__ recon,
                                                                 double* __restrict__ huber_sums) {
  extern __shared__ float s_cm[];
  float* a = s_cm;                   // [8 * 48 * 48]
  float* b = a + kCmC * 48 * 48;     // [8 * 48 * 48]
  float* s_z = b + kCmC * 48 * 48;   // [latent_dim]
  __shared__ float s_red[kCmThreads / 32];
  const int n = blockIdx.x;
  const float* xin = x + static_cast<int64_t>(n) * cin * 2304;
  // stage the input frame (read again at the end by the Huber term)
  for (int i = threadIdx.x; i < cin * 2304; i += kCmThreads) b[i] = __ldg(xin + i);
  __syncthreads();
  conv3x3(b, cin, 48, 48, W.conv0_w, W.conv0_b, kCmC, a);
  layernorm_lrelu(a, kCmC * 2304, W.ln_w[0], W.ln_b[0], s_red);
  conv4x4s2(a, 48, 48, W.down_w[0], W.down_b[0], b);
  layernorm_lrelu(b, kCmC * 576, W.ln_w[1], W.ln_b[1], s_red);
  conv4x4s2(b, 24, 24, W.down_w[1], W.down_b[1], a);
  layernorm_lrelu(a, kCmC * 144, W.ln_w[2], W.ln_b[2], s_red);
  conv4x4s2(a, 12, 12, W.down_w[2], W.down_b[2], b);
  layernorm_lrelu(b, kCmC * 36, W.ln_w[3], W.ln_b[3], s_red);
  linear(b, kCmC * 36, W.to_latent_w, W.to_latent_b, latent_dim, s_z);   // flatten order (c, h, w) = memory order
  for (int i = threadIdx.x; i < latent_dim; i += kCmThreads) z[static_cast<int64_t>(n) * latent_dim + i] = s_z[i];
  linear(s_z, latent_dim, W.to_rec_w, W.to_rec_b, kCmC * 36, a);
  convT4x4s2(a, 6, 6, W.up_w[0], W.up_b[0], b);
  layernorm_lrelu(b, kCmC * 144, W.ln_w[4], W.ln_b[4], s_red);
  convT4x4s2(b, 12, 12, W.up_w[1], W.up_b[1], a);
  layernorm_lrelu(a, kCmC * 576, W.ln_w[5], W.ln_b[5], s_red);
  convT4x4s2(a, 24, 24, W.up_w[2], W.up_b[2], b);
  layernorm_lrelu(b, kCmC * 2304, W.ln_w[6], W.ln_b[6], s_red);
  conv3x3(b, kCmC, 48, 48, W.out_w, W.out_b, cin, a);
  float hub = 0.f;
  for (int i = threadIdx.x; i < cin * 2304; i += kCmThreads) {
    const float v = a[i];
    recon[static_cast<int64_t>(n) * cin * 2304 + i] = v;
    const float d = fabsf(v - __ldg(xin + i));
    hub += d < 1.f ? 0.5f * d * d : d - 0.5f;   // nn.HuberLoss(delta = 1)
  }
  if (huber_sums != nullptr) {
    const float tot = block_sum(hub, s_red);
    if (threadIdx.x == 0) {
      atomicAdd(&huber_sums[0], static_cast<double>(tot));
      atomicAdd(&huber_sums[1], static_cast<double>(cin * 2304));
    }
  }
}

}  // namespace wfk

extern "C" int wfk_convmodel_forward(const float* x, int n, int cin, int latent_dim, const float* const* weights,
                                     float* z, float* recon, double* huber_sums, void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && weights && z && recon, "null pointer");
  WFK_REQUIRE(n > 0 && cin >= 1 && cin <= 8 && latent_dim >= 1 && latent_dim <= 4096, "bad shape");
  for (int i = 0; i < 34; ++i) WFK_REQUIRE(weights[i] != nullptr, "weights[%d] is NULL", i);
  wfk::ConvModelWeights W;
  int k = 0;
  W.conv0_w = weights[k++];
  W.conv0_b = weights[k++];
  for (int i = 0; i < 7; ++i) {
    W.ln_w[i] = weights[k++];
    W.ln_b[i] = weights[k++];
  }
  for (int i = 0; i < 3; ++i) {
    W.down_w[i] = weights[k++];
    W.down_b[i] = weights[k++];
  }
  W.to_latent_w = weights[k++];
  W.to_latent_b = weights[k++];
  W.to_rec_w = weights[k++];
  W.to_rec_b = weights[k++];
  for (int i = 0; i < 3; ++i) {
    W.up_w[i] = weights[k++];
    W.up_b[i] = weights[k++];
  }
  W.out_w = weights[k++];
  W.out_b = weights[k++];
  const size_t smem = (static_cast<size_t>(2) * wfk::kCmC * 2304 + latent_dim) * sizeof(float);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::convmodel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_lock.finished();
  }
  WFK_REQUIRE(smem <= 200 * 1024, "latent_dim too large");
  wfk::convmodel_kernel<<<n, wfk::kCmThreads, smem, static_cast<cudaStream_t>(stream)>>>(x, cin, latent_dim, W, z, recon,
                                                                                        huber_sums);
  return wfk::launched("convmodel_kernel");
}
