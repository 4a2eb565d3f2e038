// Fused skill-score pass: ONE read of (pred, target) produces every quantity calc_metrics needs
// (reference pipeline/metrics.py:86-133 makes 41 passes and ~41 + B*12 host syncs):
//   * clamp(0,1)                                              (metrics.py:92-93)
//   * hit / miss / false-alarm counts for every threshold at pool 1, 4x4-avg and 16x16-avg,
//     as exact integers, via warp ballot + popc               (metrics.py:9-16, 43-69)
//   * sum |p-t| (CRPS with one member == MAE) for the three pools, sum (p-t)^2
//                                                             (metrics.py:18-41, 77-84)
//   * per-frame max / min(target) and MSE for torchmetrics' PSNR (metrics.py:77-84)
//   * SSIM: 11x11 Gaussian (sigma 1.5) window statistics, separable, valid centres only
//                                                             (metrics.py:71-75 -> torchmetrics)
//
// Bit-exactness of the pooled counts: F.avg_pool2d sums a window sequentially in row-major order in
// fp32 and divides by k*k; the pooling lanes reproduce exactly that association order.
//
// Work decomposition (v2). The pass is nominally HBM-bound (8 B per pixel pair) but the 11x11 window costs
// 2 x 11 taps x 4 maps = 88 FMAs per pixel, so the kernel is built around the fp32 pipe:
//   * one CTA = one 32-row segment of one column strip of one frame, `nwc` column warps + 1 pool warp;
//   * a column thread owns TWO adjacent columns and walks the segment in chunks of 8 output rows:
//       V pass  -- 18 input rows come straight from global memory as 64-bit loads (256 B per warp and
//                  row), are clamped, counted (ballot/popc) when they are the chunk's own rows, and are
//                  pushed through the vertical 11-tap filter in registers with PACKED fp32x2 FMAs (the
//                  pair = the thread's two columns): 4 maps (p, t, p*p + t*t, p*t) x 8 output rows;
//                  the filtered rows go to shared memory as (mu_p, mu_t) / (E[pp+tt], E[pt]) pairs;
//       H pass  -- tasks of 1 row x 8 columns read 18 columns of those pairs (128-bit loads, rows padded
//                  so that the 8 rows of a task group hit 8 different bank groups) and apply the
//                  horizontal taps again as packed FMAs (the pair = two maps), then the SSIM formula;
//   * the pool warp is decoupled (reads global memory, never waits for the column warps): one lane per
//     4x4 / 16x16 window runs the sequential fp32 chain F.avg_pool2d runs, so a 256-add chain occupies
//     one lane of one warp instead of stalling a block.
// sigma_p^2 + sigma_t^2 enters SSIM only as a sum, so E[pp] and E[tt] are filtered as ONE map and the
// variance clamp (torchmetrics >= 1.x clamps each variance at 0, older releases do not clamp) is
// applied to the sum: the three variants differ by rounding-level amounts (< 1e-5 relative on pixels of
// flat regions; tests bound it), far inside the 1e-3 tolerance of the path.
// Per-CTA partial records are reduced by two tiny kernels in a fixed order, so float results are
// run-to-run deterministic.
#include "internal.h"

namespace wfk {

constexpr int kHalo = 5;
constexpr int kSegRows = 32;     // rows per CTA (multiple of 16: pooling windows never straddle CTAs)
constexpr int kChunkRows = 8;    // output rows per V / H pass
constexpr int kVRows = kChunkRows + 2 * kHalo;
constexpr int kMaxColWarps = 6;  // 6 x 64 = 384 columns per strip (one strip for the 384-wide VIL frames)
constexpr int kMetThreadsMax = 32 * (kMaxColWarps + 1);
constexpr int kMetWarpsMax = kMaxColWarps + 1;
constexpr int kSmemPadPx = 8;    // pad pixels in front of a filtered row (the H pass reads 5 to the left)

struct MetricsParams {
  const float* pred;
  const float* tgt;
  int h, w, frames;
  int nthr;
  int clamp01;
  int nwc;    // column warps per CTA
  int ns;     // column strips per frame
  int own;    // columns owned per strip when ns > 1 (64 * nwc - 32: 16 columns of overlap either side)
  int nseg;   // row segments per frame
  int vec2;   // 64-bit loads allowed (w even, base pointers 8-byte aligned)
  int vec4;   // 128-bit loads allowed in the pool warp (w % 4 == 0, base pointers 16-byte aligned)
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
__device__ __forceinline__ float clamp01f(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

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
// Same for the two pixels a column thread owns.
__device__ __forceinline__ void count_thresholds2(const MetricsParams& p, bool v0, bool v1, const float2 pv,
                                                  const float2 tv, int (&c)[WFK_MAX_THRESHOLDS][3]) {
#pragma unroll
  for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
    if (k < p.nthr) {
      const float th = p.thr[k];
      const unsigned bp0 = __ballot_sync(0xffffffffu, v0 && (pv.x >= th));
      const unsigned bt0 = __ballot_sync(0xffffffffu, v0 && (tv.x >= th));
      const unsigned bp1 = __ballot_sync(0xffffffffu, v1 && (pv.y >= th));
      const unsigned bt1 = __ballot_sync(0xffffffffu, v1 && (tv.y >= th));
      c[k][0] += __popc(bp0 & bt0) + __popc(bp1 & bt1);
      c[k][1] += __popc(bp0) + __popc(bp1);
      c[k][2] += __popc(bt0) + __popc(bt1);
    }
  }
}

// The pool warp: one lane per K x K window owned by this CTA; sequential row-major fp32 sum, then * 1/K^2
// (exact for powers of two) -- the association order of F.avg_pool2d.
template <int K>
__device__ __forceinline__ void pool_windows(const MetricsParams& p, const float* __restrict__ pf,
                                             const float* __restrict__ tf, int y0, int y1, int own0, int own1, int lane,
                                             int* s_counts, int* s_n, float& abs_acc) {
  const int wy0 = y0 / K, wy1 = y1 / K, wx0 = own0 / K, wx1 = own1 / K;
  const int nwx = max(wx1 - wx0, 0), nwy = max(wy1 - wy0, 0);
  const int items = nwx * nwy;
  int c[WFK_MAX_THRESHOLDS][3];
#pragma unroll
  for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) c[k][0] = c[k][1] = c[k][2] = 0;
#pragma unroll 1
  for (int base = 0; base < items; base += 32) {
    const int i = base + lane;
    const bool valid = i < items;
    float sp = 0.f, st = 0.f;
    if (valid) {
      const int wy = wy0 + i / nwx, wx = wx0 + i % nwx;
      const float* rp = pf + static_cast<int64_t>(wy * K) * p.w + wx * K;
      const float* rt = tf + static_cast<int64_t>(wy * K) * p.w + wx * K;
      if (p.vec4) {
#pragma unroll 4
        for (int rr = 0; rr < K; ++rr) {
          float4 a[K / 4], b[K / 4];
#pragma unroll
          for (int q = 0; q < K / 4; ++q) {
            a[q] = __ldg(reinterpret_cast<const float4*>(rp) + q);
            b[q] = __ldg(reinterpret_cast<const float4*>(rt) + q);
          }
#pragma unroll
          for (int q = 0; q < K / 4; ++q) {
            if (p.clamp01) {
              a[q] = make_float4(clamp01f(a[q].x), clamp01f(a[q].y), clamp01f(a[q].z), clamp01f(a[q].w));
              b[q] = make_float4(clamp01f(b[q].x), clamp01f(b[q].y), clamp01f(b[q].z), clamp01f(b[q].w));
            }
            sp = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sp, a[q].x), a[q].y), a[q].z), a[q].w);
            st = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(st, b[q].x), b[q].y), b[q].z), b[q].w);
          }
          rp += p.w;
          rt += p.w;
        }
      } else {
#pragma unroll 1
        for (int rr = 0; rr < K; ++rr) {
#pragma unroll
          for (int cc = 0; cc < K; ++cc) {
            float a = __ldg(rp + cc), b = __ldg(rt + cc);
            if (p.clamp01) {
              a = clamp01f(a);
              b = clamp01f(b);
            }
            sp = __fadd_rn(sp, a);
            st = __fadd_rn(st, b);
          }
          rp += p.w;
          rt += p.w;
        }
      }
      sp = __fmul_rn(sp, 1.0f / (K * K));
      st = __fmul_rn(st, 1.0f / (K * K));
      abs_acc += fabsf(sp - st);
    }
    count_thresholds(p, valid, sp, st, c);
  }
  if (lane == 0) {   // only this warp writes the pooled slots
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
      s_counts[k * 3 + 0] = c[k][0];
      s_counts[k * 3 + 1] = c[k][1];
      s_counts[k * 3 + 2] = c[k][2];
    }
    *s_n = items;
  }
}

__global__ void __launch_bounds__(kMetThreadsMax, 2) metrics_strip_kernel(const __grid_constant__ MetricsParams p,
                                                                         TileRec* __restrict__ recs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwc = p.nwc;
  const int ncol_thr = 32 * nwc;
  const int pitch = 64 * nwc + 2 * kSmemPadPx + 2;  // pixels; the +2 (16 B) skews consecutive rows across bank groups
  float2* VA = reinterpret_cast<float2*>(smem_raw);  // [kChunkRows][pitch] (mu_p, mu_t)
  float2* VB = VA + kChunkRows * pitch;              // [kChunkRows][pitch] (E[pp + tt], E[pt])
  int* s_counts = reinterpret_cast<int*>(VB + kChunkRows * pitch);  // [pools][thresholds][3]
  int* s_n = s_counts + WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3;     // [pools]
  float* s_wred = reinterpret_cast<float*>(s_n + 4);                // [warps][8]

  const int f = blockIdx.z, strip = blockIdx.x, seg = blockIdx.y;
  const int c0 = (p.ns == 1) ? 0 : strip * p.own - 16;
  const int own0 = (p.ns == 1) ? 0 : strip * p.own;
  const int own1 = (p.ns == 1) ? p.w : min(p.w, (strip + 1) * p.own);
  const int y_lo = seg * kSegRows, y_hi = min(p.h, y_lo + kSegRows);
  const float* __restrict__ pf = p.pred + static_cast<int64_t>(f) * p.h * p.w;
  const float* __restrict__ tf = p.tgt + static_cast<int64_t>(f) * p.h * p.w;

  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3 + 4; i += blockDim.x) s_counts[i] = 0;
  __syncthreads();

  float r_abs1 = 0.f, r_sq = 0.f, r_mx = -INFINITY, r_mn = INFINITY, r_ssim = 0.f, r_abs4 = 0.f, r_abs16 = 0.f;

  if (warp == nwc) {
    // ------------------------------------------------------------------ pool warp
    pool_windows<4>(p, pf, tf, y_lo, y_hi, own0, own1, lane, s_counts + 1 * WFK_MAX_THRESHOLDS * 3, s_n + 1, r_abs4);
    pool_windows<16>(p, pf, tf, y_lo, y_hi, own0, own1, lane, s_counts + 2 * WFK_MAX_THRESHOLDS * 3, s_n + 2, r_abs16);
  } else {
    // ------------------------------------------------------------------ column warps
    const int lc = 2 * tid;            // local column of this thread's pair
    const int gc = c0 + lc;            // global column (may be < 0 or >= w: masked)
    const bool in0 = gc >= 0 && gc < p.w, in1 = gc + 1 >= 0 && gc + 1 < p.w;
    const bool vec = p.vec2 && in0 && in1;
    const bool o0 = gc >= own0 && gc < own1, o1 = gc + 1 >= own0 && gc + 1 < own1;
    float2 g2[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) g2[k] = make_float2(p.gauss[k], p.gauss[k]);
    int cnt[WFK_MAX_THRESHOLDS][3];
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) cnt[k][0] = cnt[k][1] = cnt[k][2] = 0;
    const int x_lo = max(kHalo, own0), x_hi = min(p.w - kHalo, own1);  // SSIM centres this CTA scores

#pragma unroll 1
    for (int y0 = y_lo; y0 < y_hi; y0 += kChunkRows) {
      // ---- V pass: rows y0-5 .. y0+12 -> 8 vertically filtered rows of the 4 maps, in registers
      float2 aP[kChunkRows], aT[kChunkRows], aS[kChunkRows], aX[kChunkRows];
#pragma unroll
      for (int j = 0; j < kChunkRows; ++j) aP[j] = aT[j] = aS[j] = aX[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int dy = 0; dy < kVRows; ++dy) {
        const int y = y0 - kHalo + dy;
        float2 pv = make_float2(0.f, 0.f), tv = make_float2(0.f, 0.f);
        if (y >= 0 && y < p.h) {
          const int64_t o = static_cast<int64_t>(y) * p.w + gc;
          if (vec) {
            pv = __ldg(reinterpret_cast<const float2*>(pf + o));
            tv = __ldg(reinterpret_cast<const float2*>(tf + o));
          } else {
            if (in0) {
              pv.x = __ldg(pf + o);
              tv.x = __ldg(tf + o);
            }
            if (in1) {
              pv.y = __ldg(pf + o + 1);
              tv.y = __ldg(tf + o + 1);
            }
          }
          if (p.clamp01) {
            pv = make_float2(clamp01f(pv.x), clamp01f(pv.y));
            tv = make_float2(clamp01f(tv.x), clamp01f(tv.y));
          }
        }
        if (dy >= kHalo && dy < kHalo + kChunkRows) {
          // the chunk's own rows: pool-1 contingency counts, |d|, d^2, max / min(target)
          const bool rowok = y < p.h;
          const bool v0 = rowok && o0, v1 = rowok && o1;
          const float dx = pv.x - tv.x, dyv = pv.y - tv.y;
          if (v0) {
            r_abs1 += fabsf(dx);
            r_sq = fmaf(dx, dx, r_sq);
            r_mx = fmaxf(r_mx, tv.x);
            r_mn = fminf(r_mn, tv.x);
          }
          if (v1) {
            r_abs1 += fabsf(dyv);
            r_sq = fmaf(dyv, dyv, r_sq);
            r_mx = fmaxf(r_mx, tv.y);
            r_mn = fminf(r_mn, tv.y);
          }
          count_thresholds2(p, v0, v1, pv, tv, cnt);
        }
        const float2 ss = __ffma2_rn(tv, tv, __fmul2_rn(pv, pv));
        const float2 px = __fmul2_rn(pv, tv);
#pragma unroll
        for (int j = 0; j < kChunkRows; ++j) {
          const int k = dy - j;
          if (k >= 0 && k <= 2 * kHalo) {
            const float2 g = g2[k <= kHalo ? k : 2 * kHalo - k];
            aP[j] = __ffma2_rn(g, pv, aP[j]);
            aT[j] = __ffma2_rn(g, tv, aT[j]);
            aS[j] = __ffma2_rn(g, ss, aS[j]);
            aX[j] = __ffma2_rn(g, px, aX[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kChunkRows; ++j) {
        const int si = j * pitch + kSmemPadPx + lc;
        *reinterpret_cast<float4*>(VA + si) = make_float4(aP[j].x, aT[j].x, aP[j].y, aT[j].y);
        *reinterpret_cast<float4*>(VB + si) = make_float4(aS[j].x, aX[j].x, aS[j].y, aX[j].y);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ncol_thr) : "memory");
      // ---- H pass + SSIM: tasks of 1 row x 8 columns (consecutive lanes take consecutive ROWS: conflict-free)
#pragma unroll 1
      for (int task = tid; task < kChunkRows * 8 * nwc; task += ncol_thr) {
        const int j = task & (kChunkRows - 1), kgrp = task >> 3;
        const int y = y0 + j;
        const int gx0 = c0 + 8 * kgrp;
        if (y < kHalo || y >= p.h - kHalo || gx0 >= x_hi || gx0 + 8 <= x_lo) continue;
        float2 hA[8], hB[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) hA[o] = hB[o] = make_float2(0.f, 0.f);
        const float2* ra = VA + j * pitch + kSmemPadPx + 8 * kgrp - kHalo;  // input column i = output column o + tap - 5
        const float2* rb = VB + j * pitch + kSmemPadPx + 8 * kgrp - kHalo;
#pragma unroll
        for (int i = 0; i < 8 + 2 * kHalo; ++i) {
          // columns 1..16 come in aligned pairs (128-bit), the first and the last alone (64-bit)
          float2 a, b;
          if (i == 0 || i == 17) {
            a = ra[i];
            b = rb[i];
          } else if (i & 1) {
            const float4 a4 = *reinterpret_cast<const float4*>(ra + i);
            const float4 b4 = *reinterpret_cast<const float4*>(rb + i);
            a = make_float2(a4.x, a4.y);
            b = make_float2(b4.x, b4.y);
#pragma unroll
            for (int o = 0; o < 8; ++o) {   // the odd partner (column i + 1) is consumed here too
              const int k = i + 1 - o;
              if (k >= 0 && k <= 2 * kHalo) {
                const float2 g = g2[k <= kHalo ? k : 2 * kHalo - k];
                hA[o] = __ffma2_rn(g, make_float2(a4.z, a4.w), hA[o]);
                hB[o] = __ffma2_rn(g, make_float2(b4.z, b4.w), hB[o]);
              }
            }
          } else {
            continue;  // even i in 2..16: handled with its odd predecessor
          }
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int k = i - o;
            if (k >= 0 && k <= 2 * kHalo) {
              const float2 g = g2[k <= kHalo ? k : 2 * kHalo - k];
              hA[o] = __ffma2_rn(g, a, hA[o]);
              hB[o] = __ffma2_rn(g, b, hB[o]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int gx = gx0 + o;
          const float mu_p = hA[o].x, mu_t = hA[o].y;
          const float mu_pp = mu_p * mu_p, mu_tt = mu_t * mu_t, mu_pt = mu_p * mu_t;
          const float sig_sum = fmaxf((hB[o].x - mu_pp) - mu_tt, 0.f);
          const float sig_pt = hB[o].y - mu_pt;
          const float upper = 2.f * sig_pt + p.c2;
          const float lower = sig_sum + p.c2;
          const float val = __fdividef((2.f * mu_pt + p.c1) * upper, (mu_pp + mu_tt + p.c1) * lower);
          r_ssim += (gx >= x_lo && gx < x_hi) ? val : 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ncol_thr) : "memory");   // the next chunk overwrites VA / VB
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
        if (k < p.nthr) {
          atomicAdd(&s_counts[k * 3 + 0], cnt[k][0]);
          atomicAdd(&s_counts[k * 3 + 1], cnt[k][1]);
          atomicAdd(&s_counts[k * 3 + 2], cnt[k][2]);
        }
      }
    }
  }
  // ---- deterministic block reduction of the float partials
  r_abs1 = warp_sum(r_abs1);
  r_sq = warp_sum(r_sq);
  r_mx = warp_max(r_mx);
  r_mn = -warp_max(-r_mn);
  r_ssim = warp_sum(r_ssim);
  r_abs4 = warp_sum(r_abs4);
  r_abs16 = warp_sum(r_abs16);
  if (lane == 0) {
    float* d = s_wred + warp * 8;
    d[0] = r_abs1;
    d[1] = r_sq;
    d[2] = r_mx;
    d[3] = r_ssim;
    d[4] = r_abs4;
    d[5] = r_abs16;
    d[6] = r_mn;
  }
  __syncthreads();
  TileRec* rec = recs + (static_cast<int64_t>(f) * p.nseg + seg) * p.ns + strip;
  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3; i += blockDim.x) (&rec->counts[0][0][0])[i] = s_counts[i];
  if (tid == 0) {
    rec->n[0] = max(y_hi - y_lo, 0) * max(own1 - own0, 0);
    rec->n[1] = s_n[1];
    rec->n[2] = s_n[2];
    float r[6] = {0.f, 0.f, -INFINITY, 0.f, 0.f, 0.f};
    float mnr = INFINITY;
    for (int wi = 0; wi <= nwc; ++wi) {
      const float* d = s_wred + wi * 8;
      r[0] += d[0];
      r[1] += d[1];
      r[2] = fmaxf(r[2], d[2]);
      r[3] += d[3];
      r[4] += d[4];
      r[5] += d[5];
      mnr = fminf(mnr, d[6]);
    }
    rec->abs_sum[0] = r[0];
    rec->abs_sum[1] = r[4];
    rec->abs_sum[2] = r[5];
    rec->sq_sum = r[1];
    rec->max_t = r[2];
    rec->ssim_sum = r[3];
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

namespace wfk {
// Launch geometry of metrics_strip_kernel for an h x w frame.
struct MetricsGeom {
  int nwc, ns, own, nseg;
  size_t smem;
  int tiles() const { return ns * nseg; }
};
static MetricsGeom metrics_geom(int h, int w) {
  MetricsGeom g;
  if (w <= 64 * kMaxColWarps) {
    g.nwc = (w + 63) / 64;
    g.ns = 1;
    g.own = w;
  } else {
    g.nwc = kMaxColWarps;
    g.own = 64 * kMaxColWarps - 32;
    g.ns = (w + g.own - 1) / g.own;
  }
  g.nseg = (h + kSegRows - 1) / kSegRows;
  const size_t pitch = static_cast<size_t>(64 * g.nwc + 2 * kSmemPadPx + 2);
  g.smem = 2 * kChunkRows * pitch * sizeof(float2) + (WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3 + 4) * sizeof(int) +
           kMetWarpsMax * 8 * sizeof(float);
  return g;
}
}  // namespace wfk

extern "C" size_t wfk_metrics_workspace_bytes(int frames, int h, int w) {
  if (frames <= 0 || h <= 0 || w <= 0) return 0;
  const size_t tiles = static_cast<size_t>(wfk::metrics_geom(h, w).tiles());
  return tiles * static_cast<size_t>(frames) * sizeof(wfk::TileRec) + static_cast<size_t>(frames) * sizeof(wfk::FrameRec) + 256;
}

extern "C" int wfk_metrics(const float* pred, const float* tgt, int frames, int h, int w, const float* thresholds,
                           int n_thresholds, int clamp01, wfk_metric_partials* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  WFK_ENTER_STREAM(stream);
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
  const wfk::MetricsGeom g = wfk::metrics_geom(h, w);
  p.nwc = g.nwc;
  p.ns = g.ns;
  p.own = g.own;
  p.nseg = g.nseg;
  const uintptr_t align = reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt);
  p.vec2 = (w % 2 == 0 && (align & 7) == 0) ? 1 : 0;
  p.vec4 = (w % 4 == 0 && (align & 15) == 0) ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::metrics_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(wfk::metrics_geom(1, 64 * wfk::kMaxColWarps).smem)));
    attr_lock.finished();
  }
  wfk::metrics_strip_kernel<<<dim3(g.ns, g.nseg, frames), 32 * (g.nwc + 1), g.smem, s>>>(
      p, static_cast<wfk::TileRec*>(workspace));
  int rc = wfk::launched("metrics_strip_kernel");
  if (rc != WFK_OK) return rc;
  const size_t tile_bytes = (static_cast<size_t>(g.tiles()) * frames * sizeof(wfk::TileRec) + 255) & ~static_cast<size_t>(255);
  wfk::FrameRec* frs = reinterpret_cast<wfk::FrameRec*>(static_cast<uint8_t*>(workspace) + tile_bytes);
  wfk::metrics_frame_kernel<<<frames, 128, 0, s>>>(static_cast<const wfk::TileRec*>(workspace), g.tiles(), h, w, frs);
  rc = wfk::launched("metrics_frame_kernel");
  if (rc != WFK_OK) return rc;
  wfk::metrics_finalize_kernel<<<1, 128, 0, s>>>(frs, frames, n_thresholds, out);
  return wfk::launched("metrics_finalize_kernel");
}
