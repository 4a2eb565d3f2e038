// Latent predictor: the residual framing + nn.Linear(t_in*C -> t_out*C) + permutes of the reference
// validation_step (experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:67, 101-113) as
// ONE pass that reads/writes the [B, T, C, h, w] latent layout directly (no permute copies):
//   x[k]  = lat[b, k/C, k%C, p] - lat[b, t_in-1, k%C, p]            (train.py:104-106)
//   y[o]  = bias[o] + sum_k W[o,k] x[k]                             (train.py:108)
//   pred[b, o/C, o%C, p] = y[o] + lat[b, t_in-1, o%C, p]            (train.py:112)
// plus the pass-through target latents and the val_loss partial sum (train.py:109).
// Bandwidth-bound (25 latent frames per sequence); the 10 KB weight matrix lives in shared memory.
#include "internal.h"

namespace wfk {

constexpr int kPredThreads = 128;

__global__ void __launch_bounds__(kPredThreads) predict_linear_kernel(
    const float* __restrict__ lat, const float* __restrict__ weight, const float* __restrict__ bias, int b, int t_in,
    int t_out, int c, int hw, float* __restrict__ pred, float* __restrict__ tgt, double* __restrict__ loss_sums) {
  extern __shared__ float s_mem[];
  const int K = t_in * c, N = t_out * c;
  float* s_w = s_mem;                  // [N][K]
  float* s_b = s_w + N * K;            // [N]
  float* s_x = s_b + N;                // [K][kPredThreads]
  __shared__ float s_red[kPredThreads / 32];
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) s_w[i] = weight[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_b[i] = bias[i];
  __syncthreads();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = idx < static_cast<int64_t>(b) * hw;
  float loss = 0.f;
  if (valid) {
    const int bi = static_cast<int>(idx / hw), p = static_cast<int>(idx - static_cast<int64_t>(bi) * hw);
    const float* lb = lat + static_cast<int64_t>(bi) * (t_in + t_out) * c * hw + p;
    for (int k = 0; k < K; ++k) {
      const int ch = k % c;
      s_x[k * kPredThreads + threadIdx.x] = lb[static_cast<int64_t>(k) * hw] - lb[static_cast<int64_t>((t_in - 1) * c + ch) * hw];
    }
    for (int o = 0; o < N; ++o) {
      float acc = 0.f;
      const float* wr = s_w + o * K;
      for (int k = 0; k < K; ++k) acc = fmaf(s_x[k * kPredThreads + threadIdx.x], wr[k], acc);
      acc += s_b[o];
      const int ch = o % c;
      const float last = lb[static_cast<int64_t>((t_in - 1) * c + ch) * hw];
      const float tv = lb[static_cast<int64_t>(K + o) * hw];
      const int64_t oi = (static_cast<int64_t>(bi) * N + o) * hw + p;
      pred[oi] = acc + last;
      if (tgt != nullptr) tgt[oi] = (tv - last) + last;   // reference subtracts then re-adds the last frame
      const float d = acc - (tv - last);
      loss = fmaf(d, d, loss);
    }
  }
  if (loss_sums != nullptr) {
    for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < kPredThreads / 32; ++i) tot += s_red[i];
      atomicAdd(&loss_sums[0], tot);
      const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x;
      int64_t cnt = static_cast<int64_t>(b) * hw - first;
      cnt = cnt > kPredThreads ? kPredThreads : (cnt < 0 ? 0 : cnt);
      atomicAdd(&loss_sums[1], static_cast<double>(cnt) * N);
    }
  }
}

}  // namespace wfk

extern "C" int wfk_predict_linear(const float* lat, const float* weight, const float* bias, int b, int t_in, int t_out,
                                  int c, int hw, float* pred, float* tgt, double* loss_sums, void* stream) {
  WFK_REQUIRE_INIT();
  WFK_REQUIRE(lat && weight && bias && pred, "null pointer");
  WFK_REQUIRE(b > 0 && t_in > 0 && t_out > 0 && c > 0 && hw > 0, "empty problem");
  const int K = t_in * c, N = t_out * c;
  const size_t smem = (static_cast<size_t>(N) * K + N + static_cast<size_t>(K) * wfk::kPredThreads) * sizeof(float);
  WFK_REQUIRE(smem <= 200 * 1024, "predictor too large for shared memory (K=%d N=%d)", K, N);
  static bool attr_set = false;
  if (!attr_set) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::predict_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int64_t total = static_cast<int64_t>(b) * hw;
  const unsigned blocks = static_cast<unsigned>((total + wfk::kPredThreads - 1) / wfk::kPredThreads);
  wfk::predict_linear_kernel<<<blocks, wfk::kPredThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      lat, weight, bias, b, t_in, t_out, c, hw, pred, tgt, loss_sums);
  return wfk::launched("predict_linear_kernel");
}
