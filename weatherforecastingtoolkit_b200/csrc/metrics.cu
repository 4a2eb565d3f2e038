// Fused skill-score pass: ONE read of (pred, target) produces every quantity calc_metrics needs
// (reference pipeline/metrics.py:86-133 makes 41 passes and ~41 + B*12 host syncs):
//   * clamp(0,1)                                              (metrics.py:92-93)
//   * hit / miss / false-alarm counts for every threshold at pool 1, 4x4-avg and 16x16-avg,
//     as exact integers, via warp ballot + popc               (metrics.py:9-16, 43-69)
//   * sum |p-t| (CRPS with one member == MAE) for the three pools, sum (p-t)^2
//                                                             (metrics.py:18-41, 77-84)
//   * per-frame max(target) and MSE for torchmetrics' PSNR    (metrics.py:77-84)
//   * SSIM: 11x11 Gaussian (sigma 1.5) window statistics from shared-memory-staged tiles,
//     separable, valid centres only                           (metrics.py:71-75 -> torchmetrics)
//
// Bit-exactness of the pooled counts: F.avg_pool2d sums a window sequentially in row-major order in
// fp32 and divides by k*k; the pooling threads reproduce exactly that association order.
//
// Work decomposition: one CTA per 48x48 tile of one frame (48 = 3*16 keeps both pooling grids
// aligned); the tile plus a 5-pixel halo is staged once in shared memory. Per-tile partial records
// are reduced by a second tiny kernel in a fixed order, so float results are run-to-run deterministic.
#include "internal.h"

namespace wfk {

constexpr int kTS = 48;
constexpr int kHalo = 5;
constexpr int kRS = kTS + 2 * kHalo;  // 58
constexpr int kRSP = kRS + 1;         // padded pitch
constexpr int kMetThreads = 256;
constexpr int kMetWarps = kMetThreads / 32;

struct MetricsParams {
  const float* pred;
  const float* tgt;
  int h, w, frames;
  int nthr;
  int clamp01;
  float thr[WFK_MAX_THRESHOLDS];
  float gauss[11];
  float c1, c2;
};

struct TileRec {
  int counts[WFK_NUM_POOLS][WFK_MAX_THRESHOLDS][3];  // c_pt (tp), c_p (pred>=th), c_t (tgt>=th)
  int n[WFK_NUM_POOLS];
  int pad;
  float abs_sum[WFK_NUM_POOLS];
  float sq_sum;
  float ssim_sum;
  float max_t;
  float pad2[2];
};

struct MetSmem {
  float sp[kRS][kRSP];
  float st[kRS][kRSP];
  float hb[5][kRS][kTS];
  int counts[WFK_NUM_POOLS][WFK_MAX_THRESHOLDS][3];
  int n[WFK_NUM_POOLS];
  float wred[kMetWarps][8];  // per-warp float partials: abs1, sq, max, ssim, abs4, abs16
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Ballot/popc contingency update for one (pred, target) value per lane.
__device__ __forceinline__ void count_thresholds(const MetricsParams& p, bool valid, float pv, float tv,
                                                 int (&c)[WFK_MAX_THRESHOLDS][3]) {
#pragma unroll
  for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
    if (k < p.nthr) {
      const unsigned bp = __ballot_sync(0xffffffffu, valid && (pv >= p.thr[k]));
      const unsigned bt = __ballot_sync(0xffffffffu, valid && (tv >= p.thr[k]));
      c[k][0] += __popc(bp & bt);
      c[k][1] += __popc(bp);
      c[k][2] += __popc(bt);
    }
  }
}

// Average pooling with window K (sequential row-major fp32 sum, then * 1/K^2 -- exact for powers of
// two) over the K-aligned blocks of the tile that lie fully inside the image.
template <int K>
__device__ __forceinline__ void pool_pass(const MetricsParams& p, MetSmem& s, int x0, int y0, int pool_idx, int lane,
                                          int warp, float& abs_acc) {
  constexpr int NB = kTS / K;
  int c[WFK_MAX_THRESHOLDS][3];
#pragma unroll
  for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) c[k][0] = c[k][1] = c[k][2] = 0;
  int nvalid = 0;
  constexpr int ITERS = (NB * NB + kMetThreads - 1) / kMetThreads;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int i = it * kMetThreads + threadIdx.x;
    const int by = i / NB, bx = i - by * NB;
    const bool valid = (i < NB * NB) && (y0 + (by + 1) * K <= p.h) && (x0 + (bx + 1) * K <= p.w);
    float sp = 0.f, st = 0.f;
    if (valid) {
#pragma unroll 1
      for (int rr = 0; rr < K; ++rr) {
        const float* rp = &s.sp[kHalo + by * K + rr][kHalo + bx * K];
        const float* rt = &s.st[kHalo + by * K + rr][kHalo + bx * K];
#pragma unroll
        for (int cc = 0; cc < K; ++cc) {
          sp = __fadd_rn(sp, rp[cc]);
          st = __fadd_rn(st, rt[cc]);
        }
      }
      sp = __fmul_rn(sp, 1.0f / (K * K));
      st = __fmul_rn(st, 1.0f / (K * K));
      abs_acc += fabsf(sp - st);
    }
    nvalid += __popc(__ballot_sync(0xffffffffu, valid));
    count_thresholds(p, valid, sp, st, c);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
      if (k < p.nthr) {
        atomicAdd(&s.counts[pool_idx][k][0], c[k][0]);
        atomicAdd(&s.counts[pool_idx][k][1], c[k][1]);
        atomicAdd(&s.counts[pool_idx][k][2], c[k][2]);
      }
    }
    atomicAdd(&s.n[pool_idx], nvalid);
  }
}

__global__ void __launch_bounds__(kMetThreads, 2) metrics_tile_kernel(const __grid_constant__ MetricsParams p,
                                                                     TileRec* __restrict__ recs) {
  extern __shared__ uint8_t smem_raw[];
  MetSmem& s = *reinterpret_cast<MetSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kTS, y0 = blockIdx.y * kTS;
  const float* pf = p.pred + static_cast<int64_t>(f) * p.h * p.w;
  const float* tf = p.tgt + static_cast<int64_t>(f) * p.h * p.w;

  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3; i += kMetThreads) (&s.counts[0][0][0])[i] = 0;
  if (tid < WFK_NUM_POOLS) s.n[tid] = 0;
  // ---- stage tile + halo, clamped to [0,1] (zeros outside the image)
  for (int i = tid; i < kRS * kRS; i += kMetThreads) {
    const int r = i / kRS, c = i - r * kRS;
    const int y = y0 - kHalo + r, x = x0 - kHalo + c;
    float pv = 0.f, tv = 0.f;
    if (y >= 0 && y < p.h && x >= 0 && x < p.w) {
      pv = __ldg(pf + static_cast<int64_t>(y) * p.w + x);
      tv = __ldg(tf + static_cast<int64_t>(y) * p.w + x);
      if (p.clamp01) {
        pv = fminf(fmaxf(pv, 0.f), 1.f);
        tv = fminf(fmaxf(tv, 0.f), 1.f);
      }
    }
    s.sp[r][c] = pv;
    s.st[r][c] = tv;
  }
  __syncthreads();

  // ---- pool 1: counts, |d|, d^2, max(target) over the owned pixels
  float abs1 = 0.f, sq = 0.f, mx = -INFINITY, mn = INFINITY;
  {
    int c[WFK_MAX_THRESHOLDS][3];
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) c[k][0] = c[k][1] = c[k][2] = 0;
    int nvalid = 0;
#pragma unroll 1
    for (int i = tid; i < kTS * kTS; i += kMetThreads) {
      const int r = i / kTS, cc = i - r * kTS;
      const bool valid = (y0 + r < p.h) && (x0 + cc < p.w);
      const float pv = s.sp[r + kHalo][cc + kHalo], tv = s.st[r + kHalo][cc + kHalo];
      if (valid) {
        const float d = pv - tv;
        abs1 += fabsf(d);
        sq = fmaf(d, d, sq);
        mx = fmaxf(mx, tv);
        mn = fminf(mn, tv);
      }
      nvalid += __popc(__ballot_sync(0xffffffffu, valid));
      count_thresholds(p, valid, pv, tv, c);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
        if (k < p.nthr) {
          atomicAdd(&s.counts[0][k][0], c[k][0]);
          atomicAdd(&s.counts[0][k][1], c[k][1]);
          atomicAdd(&s.counts[0][k][2], c[k][2]);
        }
      }
      atomicAdd(&s.n[0], nvalid);
    }
  }
  // ---- pools 4 and 16
  float abs4 = 0.f, abs16 = 0.f;
  pool_pass<4>(p, s, x0, y0, 1, lane, warp, abs4);
  pool_pass<16>(p, s, x0, y0, 2, lane, warp, abs16);

  // ---- SSIM, horizontal 11-tap pass over all staged rows: 5 maps (p, t, pp, tt, pt)
  for (int i = tid; i < kRS * kTS; i += kMetThreads) {
    const int r = i / kTS, c = i - r * kTS;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float pv = s.sp[r][c + k], tv = s.st[r][c + k], g = p.gauss[k];
      const float gp = g * pv, gt = g * tv;
      a0 += gp;
      a1 += gt;
      a2 = fmaf(gp, pv, a2);
      a3 = fmaf(gt, tv, a3);
      a4 = fmaf(gp, tv, a4);
    }
    s.hb[0][r][c] = a0;
    s.hb[1][r][c] = a1;
    s.hb[2][r][c] = a2;
    s.hb[3][r][c] = a3;
    s.hb[4][r][c] = a4;
  }
  __syncthreads();
  // ---- vertical pass + SSIM map on the owned, valid window centres
  float ssim = 0.f;
  for (int i = tid; i < kTS * kTS; i += kMetThreads) {
    const int r = i / kTS, c = i - r * kTS;
    const int y = y0 + r, x = x0 + c;
    if (y >= kHalo && y < p.h - kHalo && x >= kHalo && x < p.w - kHalo) {
      float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = p.gauss[k];
#pragma unroll
        for (int q = 0; q < 5; ++q) m[q] = fmaf(g, s.hb[q][r + k][c], m[q]);
      }
      const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
      const float sig_p = fmaxf(m[2] - mu_pp, 0.f);
      const float sig_t = fmaxf(m[3] - mu_tt, 0.f);
      const float sig_pt = m[4] - mu_pt;
      const float upper = 2.f * sig_pt + p.c2;
      const float lower = sig_p + sig_t + p.c2;
      ssim += ((2.f * mu_pt + p.c1) * upper) / ((mu_pp + mu_tt + p.c1) * lower);
    }
  }
  // ---- deterministic block reduction of the float partials
  abs1 = warp_sum(abs1);
  sq = warp_sum(sq);
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  ssim = warp_sum(ssim);
  abs4 = warp_sum(abs4);
  abs16 = warp_sum(abs16);
  if (lane == 0) {
    s.wred[warp][0] = abs1;
    s.wred[warp][1] = sq;
    s.wred[warp][2] = mx;
    s.wred[warp][3] = ssim;
    s.wred[warp][4] = abs4;
    s.wred[warp][5] = abs16;
    s.wred[warp][6] = mn;
  }
  __syncthreads();
  TileRec* rec = recs + (static_cast<int64_t>(f) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3; i += kMetThreads)
    (&rec->counts[0][0][0])[i] = (&s.counts[0][0][0])[i];
  if (tid < WFK_NUM_POOLS) rec->n[tid] = s.n[tid];
  if (tid == 0) {
    float r[6] = {0.f, 0.f, -INFINITY, 0.f, 0.f, 0.f};
    for (int wi = 0; wi < kMetWarps; ++wi) {
      r[0] += s.wred[wi][0];
      r[1] += s.wred[wi][1];
      r[2] = fmaxf(r[2], s.wred[wi][2]);
      r[3] += s.wred[wi][3];
      r[4] += s.wred[wi][4];
      r[5] += s.wred[wi][5];
    }
    rec->abs_sum[0] = r[0];
    rec->abs_sum[1] = r[4];
    rec->abs_sum[2] = r[5];
    rec->sq_sum = r[1];
    rec->max_t = r[2];
    rec->ssim_sum = r[3];
    float mnr = INFINITY;
    for (int wi = 0; wi < kMetWarps; ++wi) mnr = fminf(mnr, s.wred[wi][6]);
    rec->pad2[0] = mnr;  // min(target)
  }
}

// Reduction of the per-tile records in a fixed order (deterministic float results), two levels:
// (1) one block per frame folds that frame's tiles into a FrameRec (and forms the per-frame SSIM / PSNR),
// (2) one block folds the FrameRecs in frame order into the output struct.
constexpr int kNumInt = WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3 + WFK_NUM_POOLS;  // 75

struct FrameRec {
  long long ints[kNumInt + 1];  // counts (c_pt, c_p, c_t) then n per pool
  double f[6];                  // abs1, abs4, abs16, sq, ssim_frame, psnr_frame
};

__global__ void __launch_bounds__(128) metrics_frame_kernel(const TileRec* __restrict__ recs, int tiles_per_frame,
                                                            int h, int w, FrameRec* __restrict__ frames_out) {
  __shared__ float s_f[8];
  const int f = blockIdx.x, tid = threadIdx.x;
  const TileRec* fr = recs + static_cast<int64_t>(f) * tiles_per_frame;
  FrameRec* out = frames_out + f;
  if (tid < kNumInt) {
    long long acc = 0;
    for (int t = 0; t < tiles_per_frame; ++t) {
      const int* ip = (tid < kNumInt - WFK_NUM_POOLS) ? (&fr[t].counts[0][0][0] + tid) : (&fr[t].n[0] + (tid - (kNumInt - WFK_NUM_POOLS)));
      acc += *ip;
    }
    out->ints[tid] = acc;
  } else if (tid >= 96 && tid < 96 + 7) {
    const int j = tid - 96;  // abs1, abs4, abs16, sq, ssim, max, min
    double acc = (j == 5) ? -INFINITY : ((j == 6) ? INFINITY : 0.0);
    for (int t = 0; t < tiles_per_frame; ++t) {
      const TileRec& r = fr[t];
      const float v = j == 0 ? r.abs_sum[0] : j == 1 ? r.abs_sum[1] : j == 2 ? r.abs_sum[2] : j == 3 ? r.sq_sum
                    : j == 4 ? r.ssim_sum : j == 5 ? r.max_t : r.pad2[0];
      if (j == 5) acc = fmax(acc, static_cast<double>(v));
      else if (j == 6) acc = fmin(acc, static_cast<double>(v));
      else acc += v;
    }
    if (j < 4) out->f[j] = acc;
    if (j == 4) out->f[4] = acc / (static_cast<double>(h - 2 * kHalo) * (w - 2 * kHalo));
    if (j == 3) s_f[0] = static_cast<float>(acc / (static_cast<double>(h) * w));  // mse (double kept below)
    if (j == 5) s_f[1] = static_cast<float>(acc);
    if (j == 6) s_f[2] = static_cast<float>(acc);
  }
  __syncthreads();
  if (tid == 0) {
    // torchmetrics PeakSignalNoiseRatio(data_range=None): max(target.max(), 0) - min(target.min(), 0)
    const double mse = out->f[3] / (static_cast<double>(h) * w);
    const double range = static_cast<double>(fmaxf(s_f[1], 0.f)) - static_cast<double>(fminf(s_f[2], 0.f));
    out->f[5] = 10.0 * log10(range * range / mse);
  }
}

__global__ void __launch_bounds__(128) metrics_finalize_kernel(const FrameRec* __restrict__ fr, int frames, int nthr,
                                                               wfk_metric_partials* __restrict__ out) {
  __shared__ long long s_i[kNumInt];
  __shared__ double s_d[6];
  const int tid = threadIdx.x;
  if (tid < kNumInt) {
    long long acc = 0;
    for (int f = 0; f < frames; ++f) acc += fr[f].ints[tid];
    s_i[tid] = acc;
  } else if (tid >= 96 && tid < 102) {
    double acc = 0.0;
    for (int f = 0; f < frames; ++f) acc += fr[f].f[tid - 96];
    s_d[tid - 96] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    for (int pl = 0; pl < WFK_NUM_POOLS; ++pl) {
      const long long n = s_i[kNumInt - WFK_NUM_POOLS + pl];
      for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
        const long long c_pt = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 0];
        const long long c_p = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 1];
        const long long c_t = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 2];
        const bool on = k < nthr;
        out->counts[pl][k][0] = on ? c_pt : 0;                 // tp
        out->counts[pl][k][1] = on ? c_t - c_pt : 0;           // fn  (target yes, pred no)
        out->counts[pl][k][2] = on ? c_p - c_pt : 0;           // fp  (pred yes, target no)
        out->counts[pl][k][3] = on ? n - c_p - c_t + c_pt : 0; // tn
      }
      out->n_elems[pl] = n;
      out->abs_sum[pl] = s_d[pl];
    }
    out->n_frames = frames;
    out->sq_sum = s_d[3];
    out->ssim_sum = s_d[4];
    out->psnr_sum = s_d[5];
    out->reserved[0] = out->reserved[1] = 0.0;
  }
}

}  // namespace wfk

extern "C" size_t wfk_metrics_workspace_bytes(int frames, int h, int w) {
  if (frames <= 0 || h <= 0 || w <= 0) return 0;
  const size_t tiles = static_cast<size_t>((h + wfk::kTS - 1) / wfk::kTS) * ((w + wfk::kTS - 1) / wfk::kTS);
  return tiles * static_cast<size_t>(frames) * sizeof(wfk::TileRec) + static_cast<size_t>(frames) * sizeof(wfk::FrameRec) + 256;
}

extern "C" int wfk_metrics(const float* pred, const float* tgt, int frames, int h, int w, const float* thresholds,
                           int n_thresholds, int clamp01, wfk_metric_partials* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  WFK_REQUIRE_INIT();
  WFK_REQUIRE(pred && tgt && out && workspace && thresholds, "null pointer");
  WFK_REQUIRE(frames > 0 && frames <= 65535, "frames=%d unsupported (1..65535 per call)", frames);
  WFK_REQUIRE(h >= 11 && w >= 11, "SSIM needs h, w >= 11 (got %dx%d)", h, w);
  WFK_REQUIRE(n_thresholds >= 1 && n_thresholds <= WFK_MAX_THRESHOLDS, "n_thresholds must be 1..%d", WFK_MAX_THRESHOLDS);
  WFK_REQUIRE(workspace_bytes >= wfk_metrics_workspace_bytes(frames, h, w), "workspace too small");
  wfk::MetricsParams p{};
  p.pred = pred;
  p.tgt = tgt;
  p.h = h;
  p.w = w;
  p.frames = frames;
  p.nthr = n_thresholds;
  p.clamp01 = clamp01 ? 1 : 0;
  for (int i = 0; i < n_thresholds; ++i) p.thr[i] = thresholds[i];
  // torchmetrics _gaussian(kernel_size=11, sigma=1.5) in float32
  {
    float g[11], sum = 0.f;
    for (int i = 0; i < 11; ++i) {
      const float d = static_cast<float>(i - 5) / 1.5f;
      g[i] = expf(-(d * d) / 2.f);
      sum += g[i];
    }
    for (int i = 0; i < 11; ++i) p.gauss[i] = g[i] / sum;
  }
  p.c1 = static_cast<float>(0.01 * 0.01);
  p.c2 = static_cast<float>(0.03 * 0.03);
  const int tx = (w + wfk::kTS - 1) / wfk::kTS, ty = (h + wfk::kTS - 1) / wfk::kTS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::metrics_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(wfk::MetSmem))));
    attr_set = true;
  }
  wfk::metrics_tile_kernel<<<dim3(tx, ty, frames), wfk::kMetThreads, sizeof(wfk::MetSmem), s>>>(
      p, static_cast<wfk::TileRec*>(workspace));
  int rc = wfk::launched("metrics_tile_kernel");
  if (rc != WFK_OK) return rc;
  const size_t tile_bytes = (static_cast<size_t>(tx) * ty * frames * sizeof(wfk::TileRec) + 255) & ~static_cast<size_t>(255);
  wfk::FrameRec* frs = reinterpret_cast<wfk::FrameRec*>(static_cast<uint8_t*>(workspace) + tile_bytes);
  wfk::metrics_frame_kernel<<<frames, 128, 0, s>>>(static_cast<const wfk::TileRec*>(workspace), tx * ty, h, w, frs);
  rc = wfk::launched("metrics_frame_kernel");
  if (rc != WFK_OK) return rc;
  wfk::metrics_finalize_kernel<<<1, 128, 0, s>>>(frs, frames, n_thresholds, out);
  return wfk::launched("metrics_finalize_kernel");
}
