// HBM-bound passes between the tensor-core kernels: GroupNorm apply (+SiLU), row softmax,
// fp32 -> fp16 weight conversion. All use 128-bit vector loads/stores on NHWC fp16 activations.
#include "act16.cuh"
#include "internal.h"

namespace wfk {

// y = silu?( (x - mean) * rstd * gamma + beta ), with mean / rstd from the (sum, sumsq) pairs the
// producing kernel accumulated. One block = `ppb` pixels of one frame, all channels.
template <bool BF16>
__global__ void __launch_bounds__(256) gn_apply_kernel(const uint16_t* __restrict__ x, const double* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int hw, int c, int groups,
                                                       float eps, int apply_silu, uint16_t* __restrict__ out, int ppb,
                                                       int* __restrict__ nonfinite) {
  extern __shared__ float s_ab[];  // a[c], b[c], then mean[groups], rstd[groups]
  float* s_a = s_ab;
  float* s_b = s_ab + c;
  float* s_mean = s_b + c;
  float* s_rstd = s_mean + groups;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  if (threadIdx.x < groups) {  // fp64 only once per group, not per channel
    const double cnt = static_cast<double>(cpg) * hw;
    const int g = threadIdx.x;
    const double sum = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0];
    const double sq = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1];
    if (!(isfinite(sum) && isfinite(sq))) *nonfinite = 2;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const float a = s_rstd[g] * gamma[ch];
    s_a[ch] = a;
    s_b[ch] = beta[ch] - s_mean[g] * a;
  }
  __syncthreads();
  const int vpp = c >> 3;  // 16-byte vectors per pixel
  const int p0 = blockIdx.x * ppb;
  const int npix = min(ppb, hw - p0);
  const int total = npix * vpp;
  const uint4* xin = reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(n) * hw + p0) * c);
  uint4* yout = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(n) * hw + p0) * c);
  // blockDim.x (256) is a multiple of vpp for every supported c, so a thread's channel slice is fixed:
  // its 8 scale/shift pairs live in registers and the loop is pure streaming (4 loads in flight).
  const int cv = (threadIdx.x % vpp) << 3;
  float ra[8], rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ra[j] = s_a[cv + j];
    rb[j] = s_b[cv + j];
  }
  constexpr int U = 4;
  for (int v0 = threadIdx.x; v0 < total; v0 += U * blockDim.x) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int v = v0 + k * blockDim.x;
      if (v < total) u[k] = __ldg(xin + v);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int v = v0 + k * blockDim.x;
      if (v < total) {
        uint32_t* h2 = reinterpret_cast<uint32_t*>(&u[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = A16<BF16>::unpack(h2[e]);
          f.x = fmaf(f.x, ra[2 * e], rb[2 * e]);
          f.y = fmaf(f.y, ra[2 * e + 1], rb[2 * e + 1]);
          if (apply_silu) {
            f.x = __fdividef(f.x, 1.f + __expf(-f.x));
            f.y = __fdividef(f.y, 1.f + __expf(-f.y));
          }
          h2[e] = A16<BF16>::pack(f.x, f.y);
        }
        yout[v] = u[k];
      }
    }
  }
}

// probs[r, :] = softmax(scale * scores[r, :]); one block per row, row staged in shared memory.
template <bool BF16>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ scores, int cols, float scale,
                                                           uint16_t* __restrict__ probs) {
  extern __shared__ float s_row[];
  __shared__ float s_red[8];
  const int64_t r = blockIdx.x;
  const float* src = scores + r * cols;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float v = src[i] * scale;
    s_row[i] = v;
    m = fmaxf(m, v);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float e = __expf(s_row[i] - m);
    s_row[i] = e;
    sum += e;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += s_red[i];
  const float inv = 1.f / sum;
  uint16_t* dst = probs + r * cols;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) dst[i] = A16<BF16>::pack1(s_row[i] * inv);
}

// Same for cols % 4 == 0 and cols <= 4096: the row lives in registers (up to 4 float4 per thread), one 128-bit
// load and one 64-bit store per 4 elements, two block reductions. HBM-bound: 6 bytes per score.
template <bool BF16>
__global__ void __launch_bounds__(256) softmax_rows_vec_kernel(const float* __restrict__ scores, int cols, float scale,
                                                               uint16_t* __restrict__ probs) {
  __shared__ float s_red[2][8];
  const int64_t r = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(scores + r * cols);
  const int nvec = cols >> 2;
  float4 v[4];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec) {
      float4 t = __ldcs(src + i);   // streamed once
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
      v[k] = t;
      m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
    }
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_red[0][0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[0][i]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec) {
      v[k].x = __expf(v[k].x - m); v[k].y = __expf(v[k].y - m); v[k].z = __expf(v[k].z - m); v[k].w = __expf(v[k].w - m);
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_red[1][threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += s_red[1][i];
  const float inv = 1.f / sum;
  uint2* dst = reinterpret_cast<uint2*>(probs + r * cols);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec) {
      uint2 u;
      u.x = A16<BF16>::pack(v[k].x * inv, v[k].y * inv);
      u.y = A16<BF16>::pack(v[k].z * inv, v[k].w * inv);
      dst[i] = u;
    }
  }
}

__device__ __forceinline__ float ex2_fast(float x) {   // what __expf issues after its multiply by log2(e)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// cols == 128 * NV (2304 attention tokens of a 384^2 frame: NV = 18): ONE WARP per row, the row in registers
// (NV float4 per lane, all loads in flight at once), max / sum by shuffles only. The block-per-row kernel above spends
// its time in block scheduling and two block barriers per 9 KB row (4.3 TB/s on B200 for 37 x 2304 rows).
template <bool BF16, int NV>
__global__ void __launch_bounds__(256) softmax_rows_warp_kernel(const float* __restrict__ scores, int64_t rows, float scale,
                                                                uint16_t* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int cols = 128 * NV;
  const float4* src = reinterpret_cast<const float4*>(scores + r * cols) + lane;
  float4 v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = __ldcs(src + 32 * k);   // streamed once
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NV; ++k) m = fmaxf(m, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // exp(scale * (s - max)) = 2^(s * c - max * c), c = scale * log2(e) (scale > 0: the max commutes with it)
  const float c = scale * 1.4426950408889634f;
  const float mc = -m * c;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k].x = ex2_fast(fmaf(v[k].x, c, mc));
    v[k].y = ex2_fast(fmaf(v[k].y, c, mc));
    v[k].z = ex2_fast(fmaf(v[k].z, c, mc));
    v[k].w = ex2_fast(fmaf(v[k].w, c, mc));
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  uint2* dst = reinterpret_cast<uint2*>(probs + r * cols) + lane;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    uint2 u;
    u.x = A16<BF16>::pack(v[k].x * inv, v[k].y * inv);
    u.y = A16<BF16>::pack(v[k].z * inv, v[k].w * inv);
    dst[32 * k] = u;
  }
}

template <int NV>
void launch_softmax_warp(const float* scores, int64_t rows, float scale, void* probs, int bf16, cudaStream_t s) {
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  if (bf16) softmax_rows_warp_kernel<true, NV><<<blocks, 256, 0, s>>>(scores, rows, scale, static_cast<uint16_t*>(probs));
  else softmax_rows_warp_kernel<false, NV><<<blocks, 256, 0, s>>>(scores, rows, scale, static_cast<uint16_t*>(probs));
}

// (scale, shift) per (frame, channel) of a GroupNorm, for consumers that fuse the apply step.
__global__ void gn_table_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int hw, int c, int groups, float eps,
                                float2* __restrict__ table, int* __restrict__ nonfinite) {
  // programmatic dependent launch: this grid may be scheduled while the producing convolution drains (its statistics
  // are complete and visible once the wait returns), and the consuming convolution's prologue may overlap this one
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n = blockIdx.x;
  const int cpg = c / groups;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const double cnt = static_cast<double>(cpg) * hw;
    const double s0 = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0];
    const double s1 = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1];
    if (!(isfinite(s0) && isfinite(s1))) *nonfinite = 1;   // benign race: every writer stores the same value
    const double mean = s0 / cnt;
    double var = s1 / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
    const float a = rstd * gamma[ch];
    table[static_cast<int64_t>(n) * c + ch] = make_float2(a, beta[ch] - static_cast<float>(mean) * a);
  }
}

// [n, hw, c] fp32 -> [n, c, hw] fp32 (the model-facing layout of the moments tensor)
__global__ void nhwc_to_nchw_f32_kernel(const float* __restrict__ in, int64_t total, int hw, int c,
                                        float* __restrict__ out, int* __restrict__ nonfinite) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // index into out
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const int ch = static_cast<int>((i / hw) % c);
  const int64_t n = i / (static_cast<int64_t>(hw) * c);
  const float v = in[(n * hw + p) * c + ch];
  if (!isfinite(v)) *nonfinite = 4;   // the encoder's moments: the last tensor of encode
  out[i] = v;
}

// DiagonalGaussianDistribution arithmetic in one pass over the moments tensor [n, 2*lc, hw]:
// logvar = clamp(moments[:, lc:], -30, 20); std = exp(0.5*logvar); var = exp(logvar); optionally
// sample = mean + std * noise (reference pipeline/models/autoencoderkl/distributions.py:26-42).
__global__ void __launch_bounds__(256) gaussian_posterior_kernel(const float* __restrict__ moments, int lc, int hw,
                                                                 int64_t total, float* __restrict__ logvar,
                                                                 float* __restrict__ stdv, float* __restrict__ var,
                                                                 const float* __restrict__ noise,
                                                                 float* __restrict__ sample) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // index into [n, lc, hw]
  if (i >= total) return;
  const int64_t per = static_cast<int64_t>(lc) * hw;
  const int64_t n = i / per, r = i - n * per;
  const float mean = moments[n * 2 * per + r];
  float lv = moments[n * 2 * per + per + r];
  lv = fminf(fmaxf(lv, -30.f), 20.f);
  const float sd = expf(0.5f * lv);
  if (logvar != nullptr) logvar[i] = lv;
  if (stdv != nullptr) stdv[i] = sd;
  if (var != nullptr) var[i] = expf(lv);
  if (sample != nullptr) sample[i] = mean + sd * noise[i];
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, int64_t n, __half* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

}  // namespace wfk

extern "C" int wfk_groupnorm_apply(const void* x, const double* stats, const float* gamma, const float* beta, int n,
                                   int hw, int c, int groups, float eps, int apply_silu, void* out, int bf16,
                                   void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && stats && gamma && beta && out, "null pointer");
  WFK_REQUIRE(n > 0 && hw > 0 && c > 0 && groups > 0, "empty problem");
  WFK_REQUIRE(c % 8 == 0 && c % groups == 0 && c <= 2048 && 256 % (c / 8) == 0 && groups <= 256,
              "unsupported channel count c=%d groups=%d", c, groups);
  WFK_REQUIRE(n <= 65535, "n too large");
  int ppb = 65536 / c;  // 128 KB of fp16 per block: amortises the per-block scale/shift prologue
  if (ppb < 1) ppb = 1;
  dim3 grid((hw + ppb - 1) / ppb, n);
  const size_t smem = (2 * c + 2 * groups) * sizeof(float);
  if (bf16)
    wfk::gn_apply_kernel<true><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(x), stats, gamma, beta, hw, c, groups, eps, apply_silu, static_cast<uint16_t*>(out), ppb, wfk::nonfinite_flag());
  else
    wfk::gn_apply_kernel<false><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(x), stats, gamma, beta, hw, c, groups, eps, apply_silu, static_cast<uint16_t*>(out), ppb, wfk::nonfinite_flag());
  return wfk::launched("gn_apply_kernel");
}

extern "C" int wfk_gn_table(const double* stats, const float* gamma, const float* beta, int n, int hw, int c,
                            int groups, float eps, void* table, void* stream) {
  WFK_ENTER(stream, stats);
  WFK_REQUIRE(stats && gamma && beta && table, "null pointer");
  WFK_REQUIRE(n > 0 && hw > 0 && c > 0 && groups > 0 && c % groups == 0, "bad shape");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WFK_CUDA_CHECK(cudaLaunchKernelEx(&cfg, wfk::gn_table_kernel, stats, gamma, beta, hw, c, groups, eps,
                                    static_cast<float2*>(table), wfk::nonfinite_flag()));
  return wfk::launched("gn_table_kernel");
}

extern "C" int wfk_softmax_rows(const float* scores, int64_t rows, int cols, float scale, void* probs, int bf16,
                                void* stream) {
  WFK_ENTER(stream, scores);
  WFK_REQUIRE(scores && probs, "null pointer");
  WFK_REQUIRE(rows > 0 && rows < (1ll << 31) && cols > 0 && cols <= 11264, "unsupported softmax shape");
  if (scale > 0.f && cols % 128 == 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (cols / 128) {
      case 2: wfk::launch_softmax_warp<2>(scores, rows, scale, probs, bf16, st); return wfk::launched("softmax_rows_warp_kernel");
      case 4: wfk::launch_softmax_warp<4>(scores, rows, scale, probs, bf16, st); return wfk::launched("softmax_rows_warp_kernel");
      case 8: wfk::launch_softmax_warp<8>(scores, rows, scale, probs, bf16, st); return wfk::launched("softmax_rows_warp_kernel");
      case 16: wfk::launch_softmax_warp<16>(scores, rows, scale, probs, bf16, st); return wfk::launched("softmax_rows_warp_kernel");
      case 18: wfk::launch_softmax_warp<18>(scores, rows, scale, probs, bf16, st); return wfk::launched("softmax_rows_warp_kernel");
      default: break;
    }
  }
  if (cols % 4 == 0 && cols <= 4096) {
    if (bf16)
      wfk::softmax_rows_vec_kernel<true><<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          scores, cols, scale, static_cast<uint16_t*>(probs));
    else
      wfk::softmax_rows_vec_kernel<false><<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          scores, cols, scale, static_cast<uint16_t*>(probs));
    return wfk::launched("softmax_rows_vec_kernel");
  }
  if (bf16)
    wfk::softmax_rows_kernel<true><<<static_cast<unsigned>(rows), 256, cols * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        scores, cols, scale, static_cast<uint16_t*>(probs));
  else
    wfk::softmax_rows_kernel<false><<<static_cast<unsigned>(rows), 256, cols * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        scores, cols, scale, static_cast<uint16_t*>(probs));
  return wfk::launched("softmax_rows_kernel");
}

extern "C" int wfk_gaussian_posterior(const float* moments, int n, int lc, int hw, float* logvar, float* std, float* var,
                                      const float* noise, float* sample, void* stream) {
  WFK_ENTER(stream, moments);
  WFK_REQUIRE(moments && n > 0 && lc > 0 && hw > 0, "bad argument");
  WFK_REQUIRE((sample == nullptr) == (noise == nullptr), "noise and sample must both be given or both NULL");
  const int64_t total = static_cast<int64_t>(n) * lc * hw;
  wfk::gaussian_posterior_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      moments, lc, hw, total, logvar, std, var, noise, sample);
  return wfk::launched("gaussian_posterior_kernel");
}

extern "C" int wfk_nhwc_to_nchw_f32(const float* in, int n, int hw, int c, float* out, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && out && n > 0 && hw > 0 && c > 0, "bad argument");
  const int64_t total = static_cast<int64_t>(n) * hw * c;
  wfk::nhwc_to_nchw_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, total, hw, c, out, wfk::nonfinite_flag());
  return wfk::launched("nhwc_to_nchw_f32_kernel");
}

extern "C" int wfk_f32_to_f16(const float* in, int64_t n, void* out, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && out && n > 0, "bad argument");
  const int64_t blocks = (n + 255) / 256;
  wfk::f32_to_f16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, n, static_cast<__half*>(out));
  return wfk::launched("f32_to_f16_kernel");
}
