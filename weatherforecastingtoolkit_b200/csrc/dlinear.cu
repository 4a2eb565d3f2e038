// DLinear latent predictor: series decomposition (edge-replicated moving average) + seasonal / trend
// linears of the reference DLinear modules, fused with the residual framing of their validation_step:
//   experiments/v1_experiments/pretrained_ae_dlinear_sevir/train.py:21-99 (moving_avg, series_decomp, DLinear),
//   :179-192 (inp - last, predictor, mse_loss, + last);  individual=True: ../pretrained_ae_dlinear_ind,
//   experiments/ae_s2/train.py:55-133;  the (t,c)-interleaved 52->48 form: ../pretrained_ae_dlinear_indc_indp/
//   train.py:56-99, 185-186.
// One thread owns one series ("channel" of DLinear): it reads its L input values straight from the
// [B, T, C, hw] latent layout (no permute / reshape / cat copies), keeps seasonal and trend in shared memory
// and produces the P outputs. `group` = 1: L = t_in, series = (c, pixel); `group` = C: L = t_in*C over the
// interleaved (t, c) axis, series = pixel. The reference's `individual` form is a Python loop over 9216 (2304)
// tiny nn.Linear modules; here it is the same kernel reading per-series weight rows.
// Bandwidth-bound: 25 latent frames per sequence (+ the per-series weights when individual).
#include "internal.h"

namespace wfk {

constexpr int kDlThreads = 128;

__global__ void __launch_bounds__(kDlThreads) dlinear_kernel(
    const float* __restrict__ x, int64_t x_batch_stride, const float* __restrict__ w_seas,
    const float* __restrict__ b_seas, const float* __restrict__ w_trend, const float* __restrict__ b_trend, int nb,
    int L, int P, int nc, int group, int ksize, int individual, int framed, float* __restrict__ pred,
    float* __restrict__ tgt, double* __restrict__ loss_sums) {
  extern __shared__ float s_mem[];
  float* s_seas = s_mem;                       // [L][kDlThreads]
  float* s_trend = s_seas + L * kDlThreads;    // [L][kDlThreads]
  float* s_ws = s_trend + L * kDlThreads;      // shared weights: [P][L] x2, biases [P] x2
  float* s_wt = s_ws + (individual ? 0 : P * L);
  float* s_bs = s_wt + (individual ? 0 : P * L);
  float* s_bt = s_bs + (individual ? 0 : P);
  __shared__ float s_red[kDlThreads / 32];
  if (!individual) {
    for (int i = threadIdx.x; i < P * L; i += blockDim.x) {
      s_ws[i] = w_seas[i];
      s_wt[i] = w_trend[i];
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      s_bs[i] = b_seas[i];
      s_bt[i] = b_trend[i];
    }
    __syncthreads();
  }
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = idx < static_cast<int64_t>(nb) * nc;
  float loss = 0.f;
  if (valid) {
    const int bi = static_cast<int>(idx / nc), p = static_cast<int>(idx - static_cast<int64_t>(bi) * nc);
    const float* xb = x + static_cast<int64_t>(bi) * x_batch_stride + p;
    const int last0 = L - group;  // first series position of the last input frame
    float* seas = s_seas + threadIdx.x;
    float* trend = s_trend + threadIdx.x;
    // residual w.r.t. the last input frame (train.py:183-185), staged in `seas`
    for (int s = 0; s < L; ++s) {
      float v = xb[static_cast<int64_t>(s) * nc];
      if (framed) v -= xb[static_cast<int64_t>(last0 + s % group) * nc];
      seas[s * kDlThreads] = v;
    }
    // moving_avg: AvgPool1d(k, stride 1) over the series padded with (k-1)/2 copies of its end points
    // (train.py:31-37): sequential window sum, then the division
    const int half = (ksize - 1) / 2;
    for (int s = 0; s < L; ++s) {
      float sum = 0.f;
      for (int d = -half; d <= half; ++d) {
        int q = s + d;
        q = q < 0 ? 0 : (q > L - 1 ? L - 1 : q);
        sum += seas[q * kDlThreads];
      }
      trend[s * kDlThreads] = sum / static_cast<float>(ksize);
    }
    for (int s = 0; s < L; ++s) seas[s * kDlThreads] -= trend[s * kDlThreads];  // res = x - moving_mean (:50)
    const float* ws = individual ? w_seas + static_cast<int64_t>(p) * P * L : s_ws;
    const float* wt = individual ? w_trend + static_cast<int64_t>(p) * P * L : s_wt;
    const float* bs = individual ? b_seas + static_cast<int64_t>(p) * P : s_bs;
    const float* bt = individual ? b_trend + static_cast<int64_t>(p) * P : s_bt;
    for (int j = 0; j < P; ++j) {
      float as = 0.f, at = 0.f;
      if (individual) {
        for (int s = 0; s < L; ++s) {
          as = fmaf(seas[s * kDlThreads], __ldg(ws + j * L + s), as);
          at = fmaf(trend[s * kDlThreads], __ldg(wt + j * L + s), at);
        }
        as += __ldg(bs + j);
        at += __ldg(bt + j);
      } else {
        for (int s = 0; s < L; ++s) {
          as = fmaf(seas[s * kDlThreads], ws[j * L + s], as);
          at = fmaf(trend[s * kDlThreads], wt[j * L + s], at);
        }
        as += bs[j];
        at += bt[j];
      }
      const float y = as + at;  // seasonal_output + trend_output (:96)
      const int64_t oi = (static_cast<int64_t>(bi) * P + j) * nc + p;
      if (framed) {
        const float last = xb[static_cast<int64_t>(last0 + j % group) * nc];
        const float tv = xb[static_cast<int64_t>(L + j) * nc];
        pred[oi] = y + last;
        if (tgt != nullptr) tgt[oi] = (tv - last) + last;  // the reference subtracts then re-adds the last frame
        const float d = y - (tv - last);
        loss = fmaf(d, d, loss);
      } else {
        pred[oi] = y;
      }
    }
  }
  if (loss_sums != nullptr) {
    for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < kDlThreads / 32; ++i) tot += s_red[i];
      atomicAdd(&loss_sums[0], tot);
      const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x;
      int64_t cnt = static_cast<int64_t>(nb) * nc - first;
      cnt = cnt > kDlThreads ? kDlThreads : (cnt < 0 ? 0 : cnt);
      atomicAdd(&loss_sums[1], static_cast<double>(cnt) * P);
    }
  }
}

}  // namespace wfk

extern "C" int wfk_dlinear(const float* x, int64_t x_batch_stride, const float* w_seasonal, const float* b_seasonal,
                           const float* w_trend, const float* b_trend, int nb, int seq_len, int pred_len, int channels,
                           int group, int kernel_size, int individual, int framed, float* pred, float* tgt,
                           double* loss_sums, void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && w_seasonal && b_seasonal && w_trend && b_trend && pred, "null pointer");
  WFK_REQUIRE(nb > 0 && seq_len > 0 && pred_len > 0 && channels > 0, "empty problem");
  WFK_REQUIRE(kernel_size >= 1 && (kernel_size & 1), "kernel_size=%d must be odd (the reference pads (k-1)/2 per side)",
              kernel_size);
  WFK_REQUIRE(group >= 1 && seq_len % group == 0 && pred_len % group == 0, "group=%d must divide seq_len and pred_len",
              group);
  WFK_REQUIRE(framed || (tgt == nullptr && loss_sums == nullptr), "tgt / loss need the framed (latent sequence) form");
  const size_t smem = (2 * static_cast<size_t>(seq_len) * wfk::kDlThreads +
                       (individual ? 0 : 2 * static_cast<size_t>(pred_len) * (seq_len + 1))) * sizeof(float);
  WFK_REQUIRE(smem <= 200 * 1024, "DLinear too large for shared memory (seq_len=%d pred_len=%d)", seq_len, pred_len);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::dlinear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_lock.finished();
  }
  const int64_t total = static_cast<int64_t>(nb) * channels;
  const unsigned blocks = static_cast<unsigned>((total + wfk::kDlThreads - 1) / wfk::kDlThreads);
  wfk::dlinear_kernel<<<blocks, wfk::kDlThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, x_batch_stride, w_seasonal, b_seasonal, w_trend, b_trend, nb, seq_len, pred_len, channels, group, kernel_size,
      individual, framed, pred, tgt, loss_sums);
  return wfk::launched("dlinear_kernel");
}
