// The parts of pipeline/metrics.py that calc_metrics' fixed sweep does not use but the module's public functions
// accept (reference /root/reference/pipeline/metrics.py): any pooling window (`scale`) and `pool_type='max'` in
// csi / hss / crps (:22-32, 43-50, 56-63), and ensemble forecasts (`pred.ndim == 6`: crps with n > 1 members :33-41,
// `pred.mean(dim=1)` in calc_metrics :94). Not on the hot path (the fused metrics_strip_kernel serves none / avg 4 /
// avg 16 at one member): simple one-cell-per-thread kernels, HBM-bound, same exactness rules as the fused pass --
// integer counts, fp32 compares, F.avg_pool2d's sequential row-major window sum followed by ONE division.
#include "internal.h"

namespace wfk {

constexpr int kGenThreads = 256;

__device__ __forceinline__ float gen_clamp(float v, int clamp01) { return clamp01 ? fminf(fmaxf(v, 0.f), 1.f) : v; }

// pooled value of one output cell; kind 0: none (scale 1), 1: avg, 2: max
__device__ __forceinline__ float pool_cell(const float* __restrict__ img, int w, int y0, int x0, int scale, int kind,
                                           int clamp01) {
  if (kind == 0) return gen_clamp(__ldg(img + static_cast<int64_t>(y0) * w + x0), clamp01);
  float acc = kind == 1 ? 0.f : -INFINITY;
  for (int r = 0; r < scale; ++r) {
    const float* row = img + static_cast<int64_t>(y0 + r) * w + x0;
    for (int c = 0; c < scale; ++c) {
      const float v = gen_clamp(__ldg(row + c), clamp01);
      acc = kind == 1 ? __fadd_rn(acc, v) : fmaxf(acc, v);
    }
  }
  return kind == 1 ? __fdiv_rn(acc, static_cast<float>(scale * scale)) : acc;
}

struct GenThresholds {
  float v[WFK_MAX_THRESHOLDS];
  int n;
};

// counts[k][4] (tp, fn, fp, tn) as unsigned 64-bit, sums[0] += sum |p - t| over cells, sums[1] += cells
__global__ void __launch_bounds__(kGenThreads) pooled_counts_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                                   int frames, int h, int w, int kind, int scale,
                                                                   int clamp01, const GenThresholds thr,
                                                                   unsigned long long* __restrict__ counts,
                                                                   double* __restrict__ sums) {
  const int oh = h / scale, ow = w / scale;
  const int64_t cells = static_cast<int64_t>(frames) * oh * ow;
  const int lane = threadIdx.x & 31;
  double abs_acc = 0.0;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x; base < cells; base += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const bool valid = i < cells;
    float pv = 0.f, tv = 0.f;
    if (valid) {
      const int f = static_cast<int>(i / (static_cast<int64_t>(oh) * ow));
      const int r = static_cast<int>(i - static_cast<int64_t>(f) * oh * ow);
      const int oy = r / ow, ox = r - oy * ow;
      const float* pf = pred + static_cast<int64_t>(f) * h * w;
      const float* tf = tgt + static_cast<int64_t>(f) * h * w;
      pv = pool_cell(pf, w, oy * scale, ox * scale, scale, kind, clamp01);
      tv = pool_cell(tf, w, oy * scale, ox * scale, scale, kind, clamp01);
      abs_acc += static_cast<double>(fabsf(pv - tv));
    }
    for (int k = 0; k < thr.n; ++k) {
      const unsigned bp = __ballot_sync(0xffffffffu, valid && pv >= thr.v[k]);
      const unsigned bt = __ballot_sync(0xffffffffu, valid && tv >= thr.v[k]);
      const unsigned bv = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) {
        atomicAdd(&counts[k * 4 + 0], static_cast<unsigned long long>(__popc(bp & bt)));
        atomicAdd(&counts[k * 4 + 1], static_cast<unsigned long long>(__popc(~bp & bt)));
        atomicAdd(&counts[k * 4 + 2], static_cast<unsigned long long>(__popc(bp & ~bt & bv)));
        atomicAdd(&counts[k * 4 + 3], static_cast<unsigned long long>(__popc(~bp & ~bt & bv)));
      }
    }
  }
  for (int o = 16; o; o >>= 1) abs_acc += __shfl_xor_sync(0xffffffffu, abs_acc, o);
  if (lane == 0 && abs_acc != 0.0) atomicAdd(&sums[0], abs_acc);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sums[1], static_cast<double>(cells));
}

// Gaussian CRPS of an n-member ensemble per (pooled) cell (metrics.py:33-41): mean and unbiased std over the members
// (accumulated in double like ATen's CPU std), then fp32 like the reference's tensors.
__global__ void __launch_bounds__(kGenThreads) crps_ensemble_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                                   int b, int n, int tc, int h, int w, int kind,
                                                                   int scale, int clamp01, double* __restrict__ sums) {
  const int oh = h / scale, ow = w / scale;
  const int64_t cells = static_cast<int64_t>(b) * tc * oh * ow;
  double acc = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t per_b = static_cast<int64_t>(tc) * oh * ow;
    const int bi = static_cast<int>(i / per_b);
    int64_t r = i - bi * per_b;
    const int fi = static_cast<int>(r / (static_cast<int64_t>(oh) * ow));
    r -= static_cast<int64_t>(fi) * oh * ow;
    const int oy = static_cast<int>(r / ow), ox = static_cast<int>(r - static_cast<int64_t>(oy) * ow);
    const float gt = pool_cell(tgt + (static_cast<int64_t>(bi) * tc + fi) * h * w, w, oy * scale, ox * scale, scale, kind, clamp01);
    double s = 0.0;
    for (int m = 0; m < n; ++m)
      s += pool_cell(pred + ((static_cast<int64_t>(bi) * n + m) * tc + fi) * h * w, w, oy * scale, ox * scale, scale, kind, clamp01);
    const double mean_d = s / n;
    double q = 0.0;
    for (int m = 0; m < n; ++m) {
      const double d = pool_cell(pred + ((static_cast<int64_t>(bi) * n + m) * tc + fi) * h * w, w, oy * scale, ox * scale, scale,
                                 kind, clamp01) - mean_d;
      q += d * d;
    }
    const float mean = static_cast<float>(mean_d);
    const float sd = n > 1 ? static_cast<float>(sqrt(q / (n - 1))) : 0.f;
    const float eps = 1e-10f;
    const float normed = (mean - gt + eps) / (sd + eps);
    const float cdf = 0.5f * (1.f + erff(normed * 0.70710678118654752f));
    const float pdf = expf(-0.5f * normed * normed - 0.91893853320467274f);   // log(sqrt(2 pi))
    const float val = (sd + eps) * (normed * (2.f * cdf - 1.f) + 2.f * pdf - 0.56418958354775629f);   // 1 / sqrt(pi)
    acc += static_cast<double>(val);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sums[0], acc);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sums[1], static_cast<double>(cells));
}

// out[b, inner] = (sequential fp32 sum over the n members) / n  == pred.mean(dim=1) (metrics.py:94), values clamped first
__global__ void __launch_bounds__(kGenThreads) ensemble_mean_kernel(const float* __restrict__ pred, int b, int n, int64_t inner,
                                                                   int clamp01, float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(b) * inner;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t bi = i / inner, r = i - bi * inner;
    float s = 0.f;
    for (int m = 0; m < n; ++m) s = __fadd_rn(s, gen_clamp(__ldg(pred + (bi * n + m) * inner + r), clamp01));
    out[i] = __fdiv_rn(s, static_cast<float>(n));
  }
}

static unsigned gen_grid(int64_t items) {
  const int64_t blocks = (items + kGenThreads - 1) / kGenThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace wfk

extern "C" int wfk_pooled_counts(const float* pred, const float* tgt, int frames, int h, int w, int pool_kind, int scale,
                                 const float* thresholds, int n_thresholds, int clamp01, int64_t* counts, double* sums,
                                 void* stream) {
  WFK_ENTER(stream, pred);
  WFK_REQUIRE(pred && tgt && counts && sums, "null pointer");
  WFK_REQUIRE(frames > 0 && h > 0 && w > 0, "empty problem");
  WFK_REQUIRE(pool_kind >= 0 && pool_kind <= 2 && scale >= 1 && (pool_kind != 0 || scale == 1), "bad pooling (kind %d, scale %d)",
              pool_kind, scale);
  WFK_REQUIRE(scale <= h && scale <= w, "pooling window %d larger than the %dx%d frame", scale, h, w);
  WFK_REQUIRE(n_thresholds >= 0 && n_thresholds <= WFK_MAX_THRESHOLDS && (n_thresholds == 0 || thresholds), "bad thresholds");
  wfk::GenThresholds thr{};
  thr.n = n_thresholds;
  for (int i = 0; i < n_thresholds; ++i) thr.v[i] = thresholds[i];
  const int64_t cells = static_cast<int64_t>(frames) * (h / scale) * (w / scale);
  wfk::pooled_counts_kernel<<<wfk::gen_grid(cells), wfk::kGenThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, tgt, frames, h, w, pool_kind, scale, clamp01 ? 1 : 0, thr, reinterpret_cast<unsigned long long*>(counts), sums);
  return wfk::launched("pooled_counts_kernel");
}

extern "C" int wfk_crps_ensemble(const float* pred, const float* tgt, int b, int n, int tc, int h, int w, int pool_kind,
                                 int scale, int clamp01, double* sums, void* stream) {
  WFK_ENTER(stream, pred);
  WFK_REQUIRE(pred && tgt && sums, "null pointer");
  WFK_REQUIRE(b > 0 && n > 0 && tc > 0 && h > 0 && w > 0, "empty problem");
  WFK_REQUIRE(pool_kind >= 0 && pool_kind <= 2 && scale >= 1 && (pool_kind != 0 || scale == 1) && scale <= h && scale <= w,
              "bad pooling (kind %d, scale %d)", pool_kind, scale);
  const int64_t cells = static_cast<int64_t>(b) * tc * (h / scale) * (w / scale);
  wfk::crps_ensemble_kernel<<<wfk::gen_grid(cells), wfk::kGenThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, tgt, b, n, tc, h, w, pool_kind, scale, clamp01 ? 1 : 0, sums);
  return wfk::launched("crps_ensemble_kernel");
}

extern "C" int wfk_ensemble_mean(const float* pred, int b, int n, int64_t inner, int clamp01, float* out, void* stream) {
  WFK_ENTER(stream, pred);
  WFK_REQUIRE(pred && out && b > 0 && n > 0 && inner > 0, "bad argument");
  wfk::ensemble_mean_kernel<<<wfk::gen_grid(static_cast<int64_t>(b) * inner), wfk::kGenThreads, 0,
                              static_cast<cudaStream_t>(stream)>>>(pred, b, n, inner, clamp01 ? 1 : 0, out);
  return wfk::launched("ensemble_mean_kernel");
}
