// Token-path kernels of AE_ViT_2048 (BASELINE config 4; reference
// pipeline/models/ae_vit.py:84-162). The dense contractions (patch embedding, in_proj / out_proj / feed-forward
// linears, kv_proj, unpatch) run on the tcgen05 conv-GEMM; these are the small pieces around them:
//   patchify / unpatchify   Conv2d(1, 512, 16, 16) and ConvTranspose2d(512, 1, 16, 16) as GEMMs over 16x16 patches
//                           (ae_vit.py:100, 135, 141-143, 158-159)
//   mha_small               nn.MultiheadAttention core of TransformerEncoderLayer (ae_vit.py:106-111): per (image,
//                           head) softmax(Q K^T / sqrt(dh)) V for <= 64 tokens of dh = 64, entirely in shared memory
//   cross_encode_attn       GlobalCrossEncode (ae_vit.py:4-42): one learned query per head against 64 tokens
//   layernorm_rows          post-norm LayerNorm(512, eps 1e-5) of the fp32 residual sums
//   bcast_add_rows          GlobalCrossDecode with ONE key/value token (ae_vit.py:44-82): softmax over a single key is
//                           exactly 1, so every query token receives out(v); + pos_embed (ae_vit.py:152)
#include <cuda_fp16.h>

#include "internal.h"

namespace wfk {

// img [n, 1, h, w] fp32 -> rows [n * (h/16) * (w/16)][256] fp16, element (r, s) of patch (py, px) at r*16+s
__global__ void __launch_bounds__(256) patchify16_kernel(const float* __restrict__ img, int n, int h, int w,
                                                        __half* __restrict__ rows) {
  const int pw = w >> 4, ph = h >> 4;
  const int64_t total = static_cast<int64_t>(n) * h * w;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int b = static_cast<int>(i / (static_cast<int64_t>(w) * h));
    const int64_t row = (static_cast<int64_t>(b) * ph + (y >> 4)) * pw + (x >> 4);
    rows[row * 256 + ((y & 15) << 4) + (x & 15)] = __float2half_rn(img[i]);
  }
}

// rows [n * (h/16) * (w/16)][256] fp32 -> img [n, 1, h, w] fp32
__global__ void __launch_bounds__(256) unpatchify16_kernel(const float* __restrict__ rows, int n, int h, int w,
                                                          float* __restrict__ img) {
  const int pw = w >> 4, ph = h >> 4;
  const int64_t total = static_cast<int64_t>(n) * h * w;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int b = static_cast<int>(i / (static_cast<int64_t>(w) * h));
    const int64_t row = (static_cast<int64_t>(b) * ph + (y >> 4)) * pw + (x >> 4);
    img[i] = rows[row * 256 + ((y & 15) << 4) + (x & 15)];
  }
}

constexpr int kMhaT = 64;   // max tokens per image
constexpr int kMhaD = 64;   // head dimension

// qkv [n*T][3*D] fp16 (q | k | v, heads = contiguous 64-wide slices) -> out [n*T][D] fp16. grid = (heads, n).
__global__ void __launch_bounds__(256) mha_small_kernel(const __half* __restrict__ qkv, int T, int D, float scale,
                                                       __half* __restrict__ out) {
  extern __shared__ float s_mha[];
  float (*s_q)[kMhaD + 1] = reinterpret_cast<float (*)[kMhaD + 1]>(s_mha);
  float (*s_k)[kMhaD + 1] = s_q + kMhaT;
  float (*s_v)[kMhaD + 1] = s_k + kMhaT;
  float (*s_p)[kMhaT + 1] = reinterpret_cast<float (*)[kMhaT + 1]>(s_v + kMhaT);
  const int head = blockIdx.x, b = blockIdx.y;
  const __half* base = qkv + static_cast<int64_t>(b) * T * 3 * D + head * kMhaD;
  for (int i = threadIdx.x; i < T * (kMhaD / 8); i += blockDim.x) {
    const int t = i / (kMhaD / 8), c = (i % (kMhaD / 8)) * 8;
    const __half* rowp = base + static_cast<int64_t>(t) * 3 * D + c;
    const uint4 uq = __ldg(reinterpret_cast<const uint4*>(rowp));
    const uint4 uk = __ldg(reinterpret_cast<const uint4*>(rowp + D));
    const uint4 uv = __ldg(reinterpret_cast<const uint4*>(rowp + 2 * D));
    const __half2* q2 = reinterpret_cast<const __half2*>(&uq);
    const __half2* k2 = reinterpret_cast<const __half2*>(&uk);
    const __half2* v2 = reinterpret_cast<const __half2*>(&uv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fq = __half22float2(q2[e]), fk = __half22float2(k2[e]), fv = __half22float2(v2[e]);
      s_q[t][c + 2 * e] = fq.x * scale;   // nn.MultiheadAttention scales q before the product
      s_q[t][c + 2 * e + 1] = fq.y * scale;
      s_k[t][c + 2 * e] = fk.x;
      s_k[t][c + 2 * e + 1] = fk.y;
      s_v[t][c + 2 * e] = fv.x;
      s_v[t][c + 2 * e + 1] = fv.y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int r = i / T, c = i - r * T;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < kMhaD; ++d) acc = fmaf(s_q[r][d], s_k[c][d], acc);
    s_p[r][c] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, s_p[r][c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < T; c += 32) {
      const float e = __expf(s_p[r][c] - m);
      s_p[r][c] = e;
      sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int c = lane; c < T; c += 32) s_p[r][c] *= inv;
  }
  __syncthreads();
  __half* ob = out + static_cast<int64_t>(b) * T * D + head * kMhaD;
  for (int i = threadIdx.x; i < T * kMhaD; i += blockDim.x) {
    const int r = i / kMhaD, d = i - r * kMhaD;
    float acc = 0.f;
    for (int c = 0; c < T; ++c) acc = fmaf(s_p[r][c], s_v[c][d], acc);
    ob[static_cast<int64_t>(r) * D + d] = __float2half_rn(acc);
  }
}

// q [heads][dh] fp32 (already projected and scaled); kv [n*T][2*Dl] fp16 (k | v); out [n][Dl] fp16. grid = (heads, n).
__global__ void __launch_bounds__(256) cross_encode_attn_kernel(const float* __restrict__ q, const __half* __restrict__ kv,
                                                               int T, int Dl, int dh, __half* __restrict__ out) {
  __shared__ float s_p[128];
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qh = q + head * dh;
  const __half* kb = kv + static_cast<int64_t>(b) * T * 2 * Dl + head * dh;
  for (int t = warp; t < T; t += blockDim.x >> 5) {
    float acc = 0.f;
    for (int d = lane; d < dh; d += 32) acc = fmaf(qh[d], __half2float(kb[static_cast<int64_t>(t) * 2 * Dl + d]), acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_p[t] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, s_p[t]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = __expf(s_p[t] - m);
      s_p[t] = e;
      sum += e;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int t = lane; t < T; t += 32) s_p[t] *= inv;
  }
  __syncthreads();
  const __half* vb = kb + Dl;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(s_p[t], __half2float(vb[static_cast<int64_t>(t) * 2 * Dl + d]), acc);
    out[static_cast<int64_t>(b) * Dl + head * dh + d] = __float2half_rn(acc);
  }
}

// One warp per row: y = (x - mean) / sqrt(var + eps) * gamma + beta (biased variance, like nn.LayerNorm).
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, int64_t rows, int d,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            __half* __restrict__ out_h, float* __restrict__ out_f) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + row * d;
  float sum = 0.f;
  for (int c = lane; c < d; c += 32) sum += xr[c];
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(d);
  float var = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = xr[c] - mean;
    var = fmaf(t, t, var);
  }
  for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / static_cast<float>(d) + eps);
  for (int c = lane; c < d; c += 32) {
    const float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (out_h != nullptr) out_h[row * d + c] = __float2half_rn(y);
    if (out_f != nullptr) out_f[row * d + c] = y;
  }
}

// out[b, l, :] = vec[b, :] + pos[l, :]  (fp32 in, fp16 out)
__global__ void __launch_bounds__(256) bcast_add_rows_kernel(const float* __restrict__ vec, const float* __restrict__ pos,
                                                            int n, int T, int d, __half* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(n) * T * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d);
    const int l = static_cast<int>((i / d) % T);
    const int b = static_cast<int>(i / (static_cast<int64_t>(d) * T));
    out[i] = __float2half_rn(vec[static_cast<int64_t>(b) * d + c] + pos[static_cast<int64_t>(l) * d + c]);
  }
}

inline unsigned grid_for(int64_t total, int per_block) {
  int64_t blocks = (total + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<unsigned>(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace wfk

extern "C" int wfk_patchify16(const float* img, int n, int h, int w, void* rows, void* stream) {
  WFK_ENTER(stream, img);
  WFK_REQUIRE(img && rows && n > 0 && h > 0 && w > 0 && h % 16 == 0 && w % 16 == 0, "bad argument (h, w multiples of 16)");
  const int64_t total = static_cast<int64_t>(n) * h * w;
  wfk::patchify16_kernel<<<wfk::grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, n, h, w, static_cast<__half*>(rows));
  return wfk::launched("patchify16_kernel");
}

extern "C" int wfk_unpatchify16(const float* rows, int n, int h, int w, float* img, void* stream) {
  WFK_ENTER(stream, rows);
  WFK_REQUIRE(img && rows && n > 0 && h > 0 && w > 0 && h % 16 == 0 && w % 16 == 0, "bad argument (h, w multiples of 16)");
  const int64_t total = static_cast<int64_t>(n) * h * w;
  wfk::unpatchify16_kernel<<<wfk::grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, h, w, img);
  return wfk::launched("unpatchify16_kernel");
}

extern "C" int wfk_mha_small(const void* qkv, int n, int tokens, int d_model, int heads, void* out, void* stream) {
  WFK_ENTER(stream, qkv);
  WFK_REQUIRE(qkv && out && n > 0 && n <= 65535, "bad argument");
  WFK_REQUIRE(tokens >= 1 && tokens <= wfk::kMhaT, "tokens=%d unsupported (1..%d)", tokens, wfk::kMhaT);
  WFK_REQUIRE(heads >= 1 && d_model == heads * wfk::kMhaD, "d_model=%d must be heads (%d) x 64", d_model, heads);
  constexpr size_t smem = (3 * wfk::kMhaT * (wfk::kMhaD + 1) + wfk::kMhaT * (wfk::kMhaT + 1)) * sizeof(float);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_lock.finished();
  }
  wfk::mha_small_kernel<<<dim3(heads, n), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(qkv), tokens, d_model, 0.125f, static_cast<__half*>(out));
  return wfk::launched("mha_small_kernel");
}

extern "C" int wfk_cross_encode_attn(const float* q_scaled, const void* kv, int n, int tokens, int d_latent, int heads,
                                     void* out, void* stream) {
  WFK_ENTER(stream, kv);
  WFK_REQUIRE(q_scaled && kv && out && n > 0 && n <= 65535, "bad argument");
  WFK_REQUIRE(tokens >= 1 && tokens <= 128 && heads >= 1 && d_latent % heads == 0, "bad shape");
  wfk::cross_encode_attn_kernel<<<dim3(heads, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q_scaled, static_cast<const __half*>(kv), tokens, d_latent, d_latent / heads, static_cast<__half*>(out));
  return wfk::launched("cross_encode_attn_kernel");
}

extern "C" int wfk_layernorm_rows(const float* x, int64_t rows, int d, const float* gamma, const float* beta, float eps,
                                  void* out_h, float* out_f, void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && gamma && beta && (out_h || out_f) && rows > 0 && d > 0, "bad argument");
  const int64_t blocks = (rows + 7) / 8;
  wfk::layernorm_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, d, gamma, beta, eps, static_cast<__half*>(out_h), out_f);
  return wfk::launched("layernorm_rows_kernel");
}

extern "C" int wfk_bcast_add_rows(const float* vec, const float* pos, int n, int tokens, int d, void* out, void* stream) {
  WFK_ENTER(stream, vec);
  WFK_REQUIRE(vec && pos && out && n > 0 && tokens > 0 && d > 0, "bad argument");
  const int64_t total = static_cast<int64_t>(n) * tokens * d;
  wfk::bcast_add_rows_kernel<<<wfk::grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      vec, pos, n, tokens, d, static_cast<__half*>(out));
  return wfk::launched("bcast_add_rows_kernel");
}
