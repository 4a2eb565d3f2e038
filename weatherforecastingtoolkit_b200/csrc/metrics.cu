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
// Work decomposition (v3). The pass is nominally HBM-bound (8 B per pixel pair) but the 11x11 window costs
// 2 x 11 taps x 4 maps = 88 FMAs per pixel, so the kernel is built around the fp32 pipe (measured on B200,
// scripts/probes/fp32_rate.cu: 97 FFMA lanes/clk/SM, 64 FFMA2 lanes/clk/SM = 128 FMAs/clk/SM):
//   * one CTA = one 32-row segment of one column strip of one frame, `nwc` warps of 64 columns;
//   * a thread owns TWO adjacent columns and walks the segment in chunks of 8 output rows:
//       V pass  -- 18 input rows come straight from global memory as 64-bit loads (256 B per warp and
//                  row), are clamped and pushed through the vertical 11-tap filter in registers with
//                  PACKED fp32x2 FMAs (the pair = the thread's two columns): 4 maps (p, t, p*p + t*t, p*t)
//                  x 8 output rows; the filtered rows go to shared memory as (mu_p, mu_t) / (E[pp+tt],
//                  E[pt]) pairs. The chunk's own 8 rows are also (a) counted against the thresholds as
//                  per-thread BIT MASKS (one predicated OR per compare; 3 popc per threshold and chunk
//                  instead of a ballot + popc per row: POPC issues at 1/8 rate), (b) written to shared memory
//                  as (p, t) pairs for the pooling tasks;
//       tasks   -- after one block barrier the warps pull warp-sized tasks from a shared queue:
//                  * pool 16: one lane per 16x16 window continues the sequential fp32 chain F.avg_pool2d
//                    runs (8 rows of it per chunk, the running sums carried in shared memory);
//                  * pool 4: one lane per 4x4 window;
//                  * H pass: 1 row x 16 columns per lane, 26 columns of the filtered pairs read with 128-bit
//                    loads (rows padded so that the 8 rows of a lane group hit 8 different bank groups), the
//                    horizontal taps again as packed FMAs (the pair = two maps), then the SSIM formula.
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
constexpr int kChunkRows = 8;    // output rows per V pass / task phase
constexpr int kVRows = kChunkRows + 2 * kHalo;
constexpr int kMaxColWarps = 6;  // 6 x 64 = 384 columns per strip (one strip for the 384-wide VIL frames)
constexpr int kMetThreadsMax = 32 * kMaxColWarps;
constexpr int kSmemPadPx = 8;    // pad pixels in front of a filtered row (the H pass reads 5 to the left)
constexpr int kHCols = 16;       // output columns per H-pass lane task
constexpr int kWinBytes = 16 * 8 + 16;  // raw (p, t) pairs of one 16-pixel window row + 16 B skew (conflict-free pool-16 reads)

struct MetricsParams {
  const float* pred;
  const float* tgt;
  int h, w, frames;
  int nthr;
  int clamp01;
  int nwc;    // warps (64 columns each) per CTA
  int ns;     // column strips per frame
  int own;    // columns owned per strip when ns > 1 (64 * nwc - 32: 16 columns of overlap either side)
  int nseg;   // row segments per frame
  int vec2;   // 64-bit loads allowed (w even, base pointers 8-byte aligned)
  int perm[WFK_MAX_THRESHOLDS];  // thr[] is sorted ascending; perm[k] = position of thr[k] in the caller's list
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

// Ballot/popc contingency update for one (pred, target) value per lane (pooled cells: a few warp tasks per chunk),
// accumulated straight into the CTA's shared counters by lane 0.
__device__ __forceinline__ void count_thresholds_smem(const MetricsParams& p, bool valid, float pv, float tv, int lane,
                                                      int* s_cnt) {
#pragma unroll
  for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
    if (k < p.nthr) {
      const unsigned bp = __ballot_sync(0xffffffffu, valid && (pv >= p.thr[k]));
      const unsigned bt = __ballot_sync(0xffffffffu, valid && (tv >= p.thr[k]));
      if (lane == 0 && (bp | bt)) {
        atomicAdd(&s_cnt[k * 3 + 0], __popc(bp & bt));
        atomicAdd(&s_cnt[k * 3 + 1], __popc(bp));
        atomicAdd(&s_cnt[k * 3 + 2], __popc(bt));
      }
    }
  }
}

template <bool VEC2>
__global__ void __launch_bounds__(kMetThreadsMax, 2) metrics_strip_kernel(const __grid_constant__ MetricsParams p,
                                                                         TileRec* __restrict__ recs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwc = p.nwc;
  const int nthreads = 32 * nwc;
  const int pitch = 64 * nwc + 2 * kSmemPadPx + 2;  // pixels; the +2 (16 B) skews consecutive rows across bank groups
  const int rpitch = 4 * nwc * kWinBytes;           // bytes per raw row
  float2* VA = reinterpret_cast<float2*>(smem_raw);  // [kChunkRows][pitch] (mu_p, mu_t)
  float2* VB = VA + kChunkRows * pitch;              // [kChunkRows][pitch] (E[pp + tt], E[pt])
  uint8_t* RAW = reinterpret_cast<uint8_t*>(VB + kChunkRows * pitch);   // [kChunkRows][4 nwc windows][144 B]
  int* s_counts = reinterpret_cast<int*>(RAW + kChunkRows * rpitch);    // [pools][thresholds][3]
  int* s_n = s_counts + WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3;         // [pools] (+ task counter in s_n[3])
  float* s_carry = reinterpret_cast<float*>(s_n + 4);                   // [4 nwc windows][2] running pool-16 sums
  float* s_wred = s_carry + 2 * 4 * kMaxColWarps;                       // [warps][8]
  // Task results are accumulated per ITEM (an item is served by whichever warp pulls its task, so per-thread partials
  // would group the floats differently from run to run): [H items 32 nwc | pool-4 items 32 nwc | pool-16 windows 32]
  float* s_item = s_wred + 8 * kMaxColWarps;

  const int f = blockIdx.z, strip = blockIdx.x, seg = blockIdx.y;
  const int c0 = (p.ns == 1) ? 0 : strip * p.own - 16;
  const int own0 = (p.ns == 1) ? 0 : strip * p.own;
  const int own1 = (p.ns == 1) ? p.w : min(p.w, (strip + 1) * p.own);
  const int y_lo = seg * kSegRows, y_hi = min(p.h, y_lo + kSegRows);
  const float* __restrict__ pf = p.pred + static_cast<int64_t>(f) * p.h * p.w;
  const float* __restrict__ tf = p.tgt + static_cast<int64_t>(f) * p.h * p.w;

  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3 + 4; i += nthreads) s_counts[i] = 0;
  for (int i = tid; i < 64 * nwc + 32; i += nthreads) s_item[i] = 0.f;
  __syncthreads();

  float r_abs1 = 0.f, r_sq = 0.f, r_mx = -INFINITY, r_mn = INFINITY, r_ssim = 0.f, r_abs4 = 0.f, r_abs16 = 0.f;
  int n4 = 0, n16 = 0;   // pooled cells counted by this lane

  const int lc = 2 * tid;            // local column of this thread's pair
  const int gc = c0 + lc;            // global column (may be < 0 or >= w: masked)
  const bool in0 = gc >= 0 && gc < p.w, in1 = gc + 1 >= 0 && gc + 1 < p.w;
  // clamped (always in-bounds, even) column for the branch-free loads; values of masked columns are zeroed afterwards
  const int gcl = VEC2 ? min(max(gc, 0), p.w - 2) : min(max(gc, 0), p.w - 1);
  const int gcl1 = min(max(gc + 1, 0), p.w - 1);
  const bool o0 = gc >= own0 && gc < own1, o1 = gc + 1 >= own0 && gc + 1 < own1;
  float2 g2[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) g2[k] = make_float2(p.gauss[k], p.gauss[k]);
  const int x_lo = max(kHalo, own0), x_hi = min(p.w - kHalo, own1);  // SSIM centres this CTA scores
  uint8_t* raw_mine = RAW + (lc >> 4) * kWinBytes + (lc & 15) * 8;

#pragma unroll 1
  for (int y0 = y_lo; y0 < y_hi; y0 += kChunkRows) {
    // ---- V pass: rows y0-5 .. y0+12 -> 8 vertically filtered rows of the 4 maps, in registers
    float2 aP[kChunkRows], aT[kChunkRows], aS[kChunkRows], aX[kChunkRows];
#pragma unroll
    for (int j = 0; j < kChunkRows; ++j) aP[j] = aT[j] = aS[j] = aX[j] = make_float2(0.f, 0.f);
    unsigned mp[WFK_MAX_THRESHOLDS], mt[WFK_MAX_THRESHOLDS];   // bit 2j + e: (row j, column e) >= threshold
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) mp[k] = mt[k] = 0u;
    if (tid == 0) s_n[3] = 0;   // task queue of this chunk (read after the barrier below)
    // Loads are issued in batches of 6 rows, two batches in flight (branch-free: clamped addresses, out-of-image
    // values zeroed afterwards). Loading a row right where it is consumed made every row pay a full global-memory
    // round trip: the V pass was latency-bound at 1/4 of its instruction-issue time.
    constexpr int kBatch = 6;
    float2 bp[2][kBatch], bt[2][kBatch];
    auto load_batch = [&](const int b0, float2 (&dp)[kBatch], float2 (&dt)[kBatch]) {
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        // 32-bit element offsets inside the frame (h * w < 2^31); rows clamped into the image
        const int y = min(max(y0 - kHalo + b0 + i, 0), p.h - 1);
        const unsigned row = static_cast<unsigned>(y) * static_cast<unsigned>(p.w);
        if (VEC2) {
          dp[i] = __ldg(reinterpret_cast<const float2*>(pf + (row + gcl)));
          dt[i] = __ldg(reinterpret_cast<const float2*>(tf + (row + gcl)));
        } else {
          dp[i] = make_float2(__ldg(pf + (row + gcl)), __ldg(pf + (row + gcl1)));
          dt[i] = make_float2(__ldg(tf + (row + gcl)), __ldg(tf + (row + gcl1)));
        }
      }
    };
    auto compute_batch = [&](const int b0, const float2 (&sp)[kBatch], const float2 (&st)[kBatch]) {
#pragma unroll
      for (int i = 0; i < kBatch; ++i) {
        const int dy = b0 + i;
        const int y = y0 - kHalo + dy;
        // Rows / columns outside the image carry the clamped-address values of the border: they only ever reach
        // SSIM centres and pooling windows that are not scored, and the pool-1 statistics below mask them explicitly.
        float2 pv = sp[i], tv = st[i];
        if (p.clamp01) {
          pv = make_float2(clamp01f(pv.x), clamp01f(pv.y));
          tv = make_float2(clamp01f(tv.x), clamp01f(tv.y));
        }
        if (dy >= kHalo && dy < kHalo + kChunkRows) {
          // the chunk's own rows: raw pairs for the pooling tasks, pool-1 contingency masks, |d|, d^2, max / min(target)
          const int j = dy - kHalo;
          *reinterpret_cast<float4*>(raw_mine + j * rpitch) = make_float4(pv.x, tv.x, pv.y, tv.y);
          const bool rowok = y < p.h;
          const bool v0 = rowok && o0, v1 = rowok && o1;
          const float dx = pv.x - tv.x, dyv = pv.y - tv.y;
          r_abs1 += (v0 ? fabsf(dx) : 0.f) + (v1 ? fabsf(dyv) : 0.f);
          r_sq = fmaf(v0 ? dx : 0.f, dx, r_sq);
          r_sq = fmaf(v1 ? dyv : 0.f, dyv, r_sq);
          // invalid pixels compare as -inf (and as +inf for the minimum)
          const float c_px = v0 ? pv.x : -INFINITY, c_py = v1 ? pv.y : -INFINITY;
          const float c_tx = v0 ? tv.x : -INFINITY, c_ty = v1 ? tv.y : -INFINITY;
          r_mx = fmaxf(r_mx, fmaxf(c_tx, c_ty));
          r_mn = fminf(r_mn, fminf(v0 ? tv.x : INFINITY, v1 ? tv.y : INFINITY));
          // thresholds are ascending (sorted by the host wrapper): once no pixel of the warp's 64 reaches threshold k,
          // none reaches a later one -- VIL fields are mostly below the first threshold
          const float c_max = fmaxf(fmaxf(c_px, c_py), fmaxf(c_tx, c_ty));
#pragma unroll
          for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
            if (k >= p.nthr) break;
            const float th = p.thr[k];
            if (!__any_sync(0xffffffffu, c_max >= th)) break;
            if (c_px >= th) mp[k] |= 1u << (2 * j);
            if (c_py >= th) mp[k] |= 1u << (2 * j + 1);
            if (c_tx >= th) mt[k] |= 1u << (2 * j);
            if (c_ty >= th) mt[k] |= 1u << (2 * j + 1);
          }
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
    };
    static_assert(kVRows == 3 * kBatch, "three batches of rows per chunk");
    load_batch(0, bp[0], bt[0]);
    load_batch(kBatch, bp[1], bt[1]);
    compute_batch(0, bp[0], bt[0]);
    load_batch(2 * kBatch, bp[0], bt[0]);
    compute_batch(kBatch, bp[1], bt[1]);
    compute_batch(2 * kBatch, bp[0], bt[0]);
#pragma unroll
    for (int j = 0; j < kChunkRows; ++j) {
      const int si = j * pitch + kSmemPadPx + lc;
      *reinterpret_cast<float4*>(VA + si) = make_float4(aP[j].x, aT[j].x, aP[j].y, aT[j].y);
      *reinterpret_cast<float4*>(VB + si) = make_float4(aS[j].x, aX[j].x, aS[j].y, aX[j].y);
    }
    // pool-1 counts of the chunk: 3 popc + 3 warp-wide integer reductions per threshold
#pragma unroll
    for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
      if (k < p.nthr) {
        const unsigned any = __reduce_or_sync(0xffffffffu, mp[k] | mt[k]);
        if (any) {
          const int c_pt = __reduce_add_sync(0xffffffffu, __popc(mp[k] & mt[k]));
          const int c_p = __reduce_add_sync(0xffffffffu, __popc(mp[k]));
          const int c_t = __reduce_add_sync(0xffffffffu, __popc(mt[k]));
          if (lane == 0) {
            atomicAdd(&s_counts[k * 3 + 0], c_pt);
            atomicAdd(&s_counts[k * 3 + 1], c_p);
            atomicAdd(&s_counts[k * 3 + 2], c_t);
          }
        }
      }
    }
    __syncthreads();
    // ---- task phase: warp-sized tasks from a shared queue: [pool 16] [H pass x nwc] [pool 4 x nwc]
    const int n_tasks = 1 + 2 * nwc;
    for (;;) {
      int task = 0;
      if (lane == 0) task = atomicAdd(&s_n[3], 1);
      task = __shfl_sync(0xffffffffu, task, 0);
      if (task >= n_tasks) break;
      if (task == 0) {
        // ---- pool 16: lane = window column; 8 more rows of the sequential row-major fp32 sum (F.avg_pool2d's order)
        const int gx0 = c0 + 16 * lane;
        const bool win = lane < 4 * nwc && gx0 >= own0 && gx0 < own1 && gx0 + 16 <= p.w && (y0 & ~15) + 16 <= p.h;
        const bool second = ((y0 - y_lo) >> 3) & 1;
        float sp = 0.f, st = 0.f;
        if (win) {
          if (second) {
            sp = s_carry[2 * lane];
            st = s_carry[2 * lane + 1];
          }
          const uint8_t* wr = RAW + lane * kWinBytes;
#pragma unroll 2
          for (int r = 0; r < kChunkRows; ++r) {
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(wr + r * rpitch + 16 * q);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              sp = __fadd_rn(__fadd_rn(sp, v[q].x), v[q].z);
              st = __fadd_rn(__fadd_rn(st, v[q].y), v[q].w);
            }
          }
          if (!second) {
            s_carry[2 * lane] = sp;
            s_carry[2 * lane + 1] = st;
          } else {
            sp = __fmul_rn(sp, 1.0f / 256.0f);
            st = __fmul_rn(st, 1.0f / 256.0f);
            s_item[64 * nwc + lane] += fabsf(sp - st);
            ++n16;
          }
        }
        if (second) count_thresholds_smem(p, win, sp, st, lane, s_counts + 2 * WFK_MAX_THRESHOLDS * 3);
      } else if (task <= nwc) {
        // ---- H pass + SSIM: lane = (row j, 16-column group); consecutive lanes take consecutive ROWS (conflict-free)
        const int item = (task - 1) * 32 + lane;
        const int j = item & (kChunkRows - 1), kgrp = item >> 3;
        const int y = y0 + j;
        const int gx0 = c0 + kHCols * kgrp;
        if (y >= kHalo && y < p.h - kHalo && gx0 < x_hi && gx0 + kHCols > x_lo) {
          float2 hA[kHCols], hB[kHCols];
#pragma unroll
          for (int o = 0; o < kHCols; ++o) hA[o] = hB[o] = make_float2(0.f, 0.f);
          const float2* ra = VA + j * pitch + kSmemPadPx + kHCols * kgrp - kHalo;  // input column i = output column o + tap - 5
          const float2* rb = VB + j * pitch + kSmemPadPx + kHCols * kgrp - kHalo;
          // input columns 0 .. 25: 0 and 25 alone (64-bit), (1,2) (3,4) ... (23,24) as aligned pairs (128-bit)
#pragma unroll
          for (int i = 0; i < kHCols + 2 * kHalo; ++i) {
            float2 a, b, a_n = make_float2(0.f, 0.f), b_n = make_float2(0.f, 0.f);
            bool pair = false;
            if (i == 0 || i == kHCols + 2 * kHalo - 1) {
              a = ra[i];
              b = rb[i];
            } else if (i & 1) {
              const float4 a4 = *reinterpret_cast<const float4*>(ra + i);
              const float4 b4 = *reinterpret_cast<const float4*>(rb + i);
              a = make_float2(a4.x, a4.y);
              b = make_float2(b4.x, b4.y);
              a_n = make_float2(a4.z, a4.w);
              b_n = make_float2(b4.z, b4.w);
              pair = true;
            } else {
              continue;  // even i in 2 .. 24: consumed with its odd predecessor
            }
#pragma unroll
            for (int o = 0; o < kHCols; ++o) {
              const int k = i - o;
              if (k >= 0 && k <= 2 * kHalo) {
                const float2 g = g2[k <= kHalo ? k : 2 * kHalo - k];
                hA[o] = __ffma2_rn(g, a, hA[o]);
                hB[o] = __ffma2_rn(g, b, hB[o]);
              }
              const int k2 = i + 1 - o;
              if (pair && k2 >= 0 && k2 <= 2 * kHalo) {
                const float2 g = g2[k2 <= kHalo ? k2 : 2 * kHalo - k2];
                hA[o] = __ffma2_rn(g, a_n, hA[o]);
                hB[o] = __ffma2_rn(g, b_n, hB[o]);
              }
            }
          }
          float acc = 0.f;
#pragma unroll
          for (int o = 0; o < kHCols; ++o) {
            const int gx = gx0 + o;
            const float mu_p = hA[o].x, mu_t = hA[o].y;
            const float mu_pp = mu_p * mu_p, mu_tt = mu_t * mu_t, mu_pt = mu_p * mu_t;
            const float sig_sum = fmaxf((hB[o].x - mu_pp) - mu_tt, 0.f);
            const float sig_pt = hB[o].y - mu_pt;
            const float upper = 2.f * sig_pt + p.c2;
            const float lower = sig_sum + p.c2;
            const float val = __fdividef((2.f * mu_pt + p.c1) * upper, (mu_pp + mu_tt + p.c1) * lower);
            acc += (gx >= x_lo && gx < x_hi) ? val : 0.f;
          }
          s_item[item] += acc;
        }
      } else {
        // ---- pool 4: lane = one 4x4 window of the chunk's two window rows
        const int item = (task - 1 - nwc) * 32 + lane;
        const int per_row = 16 * nwc;
        const int wrow = item / per_row, w4 = item - wrow * per_row;
        const int lx = 4 * w4, gx0 = c0 + lx;
        const bool win = wrow < 2 && gx0 >= own0 && gx0 < own1 && gx0 + 4 <= p.w && y0 + 4 * wrow + 4 <= p.h;
        float sp = 0.f, st = 0.f;
        if (win) {
          const uint8_t* wr = RAW + (4 * wrow) * rpitch + (lx >> 4) * kWinBytes + (lx & 15) * 8;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 v0 = *reinterpret_cast<const float4*>(wr + r * rpitch);
            const float4 v1 = *reinterpret_cast<const float4*>(wr + r * rpitch + 16);
            sp = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sp, v0.x), v0.z), v1.x), v1.z);
            st = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(st, v0.y), v0.w), v1.y), v1.w);
          }
          sp = __fmul_rn(sp, 1.0f / 16.0f);
          st = __fmul_rn(st, 1.0f / 16.0f);
          s_item[32 * nwc + item] += fabsf(sp - st);
          ++n4;
        }
        count_thresholds_smem(p, win, sp, st, lane, s_counts + 1 * WFK_MAX_THRESHOLDS * 3);
      }
    }
    __syncthreads();   // the next chunk overwrites VA / VB / RAW and resets the queue
  }
  // ---- deterministic block reduction of the float partials (items: thread i folds items i, i + nthreads, ...)
  for (int i = tid; i < 32 * nwc; i += nthreads) {
    r_ssim += s_item[i];
    r_abs4 += s_item[32 * nwc + i];
  }
  if (tid < 32) r_abs16 += s_item[64 * nwc + tid];
  r_abs1 = warp_sum(r_abs1);
  r_sq = warp_sum(r_sq);
  r_mx = warp_max(r_mx);
  r_mn = -warp_max(-r_mn);
  r_ssim = warp_sum(r_ssim);
  r_abs4 = warp_sum(r_abs4);
  r_abs16 = warp_sum(r_abs16);
  n4 = __reduce_add_sync(0xffffffffu, n4);
  n16 = __reduce_add_sync(0xffffffffu, n16);
  if (lane == 0) {
    float* d = s_wred + warp * 8;
    d[0] = r_abs1;
    d[1] = r_sq;
    d[2] = r_mx;
    d[3] = r_ssim;
    d[4] = r_abs4;
    d[5] = r_abs16;
    d[6] = r_mn;
    atomicAdd(&s_n[1], n4);
    atomicAdd(&s_n[2], n16);
  }
  __syncthreads();
  TileRec* rec = recs + (static_cast<int64_t>(f) * p.nseg + seg) * p.ns + strip;
  for (int i = tid; i < WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3; i += nthreads) (&rec->counts[0][0][0])[i] = s_counts[i];
  if (tid == 0) {
    rec->n[0] = max(y_hi - y_lo, 0) * max(own1 - own0, 0);
    rec->n[1] = s_n[1];
    rec->n[2] = s_n[2];
    float r[6] = {0.f, 0.f, -INFINITY, 0.f, 0.f, 0.f};
    float mnr = INFINITY;
    for (int wi = 0; wi < nwc; ++wi) {
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
#pragma unroll 4
    for (int t = 0; t < tiles_per_frame; ++t) {
      const int* ip = (tid < kNumInt - WFK_NUM_POOLS) ? (&fr[t].counts[0][0][0] + tid) : (&fr[t].n[0] + (tid - (kNumInt - WFK_NUM_POOLS)));
      acc += *ip;
    }
    out->ints[tid] = acc;
  } else if (tid >= 96 && tid < 96 + 7) {
    const int j = tid - 96;  // abs1, abs4, abs16, sq, ssim, max, min
    double acc = (j == 5) ? -INFINITY : ((j == 6) ? INFINITY : 0.0);
#pragma unroll 4
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

// One block of 32 warps; output o (75 integer words + 6 doubles) belongs to warp o % 32. A lane sums frames
// lane, lane + 32, ... in order and the warp folds the 32 partials with a fixed shuffle tree: deterministic, and 12
// loads per lane instead of a serial chain of `frames` dependent L2 round trips in one thread (~0.4 us each).
struct ThrPerm {
  int v[WFK_MAX_THRESHOLDS];
};
__global__ void __launch_bounds__(1024) metrics_finalize_kernel(const FrameRec* __restrict__ fr, int frames, int nthr,
                                                                const ThrPerm perm, wfk_metric_partials* __restrict__ out) {
  __shared__ long long s_i[kNumInt];
  __shared__ double s_d[6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = warp; o < kNumInt + 6; o += 32) {
    if (o < kNumInt) {
      long long acc = 0;
#pragma unroll 4
      for (int f = lane; f < frames; f += 32) acc += fr[f].ints[o];
#pragma unroll
      for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) s_i[o] = acc;
    } else {
      double acc = 0.0;
#pragma unroll 4
      for (int f = lane; f < frames; f += 32) acc += fr[f].f[o - kNumInt];
#pragma unroll
      for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) s_d[o - kNumInt] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int pl = 0; pl < WFK_NUM_POOLS; ++pl) {
      const long long n = s_i[kNumInt - WFK_NUM_POOLS + pl];
      for (int k = 0; k < WFK_MAX_THRESHOLDS; ++k) {
        const long long c_pt = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 0];
        const long long c_p = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 1];
        const long long c_t = s_i[(pl * WFK_MAX_THRESHOLDS + k) * 3 + 2];
        const bool on = k < nthr;
        const int ko = on ? perm.v[k] : k;                      // position in the caller's threshold list
        out->counts[pl][ko][0] = on ? c_pt : 0;                 // tp
        out->counts[pl][ko][1] = on ? c_t - c_pt : 0;           // fn  (target yes, pred no)
        out->counts[pl][ko][2] = on ? c_p - c_pt : 0;           // fp  (pred yes, target no)
        out->counts[pl][ko][3] = on ? n - c_p - c_t + c_pt : 0; // tn
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
  g.smem = 2 * kChunkRows * pitch * sizeof(float2) + static_cast<size_t>(kChunkRows) * 4 * g.nwc * kWinBytes +
           (WFK_NUM_POOLS * WFK_MAX_THRESHOLDS * 3 + 4) * sizeof(int) + 2 * 4 * kMaxColWarps * sizeof(float) +
           kMaxColWarps * 8 * sizeof(float) + (64 * kMaxColWarps + 32) * sizeof(float);
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
  WFK_ENTER(stream, pred);
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
  // the kernel wants ascending thresholds (early exit of the compare loop); results are un-permuted by finalize
  for (int i = 0; i < WFK_MAX_THRESHOLDS; ++i) p.perm[i] = i;
  for (int i = 1; i < n_thresholds; ++i)
    for (int j = i; j > 0 && p.thr[j] < p.thr[j - 1]; --j) {
      const float tf_ = p.thr[j];
      p.thr[j] = p.thr[j - 1];
      p.thr[j - 1] = tf_;
      const int ti_ = p.perm[j];
      p.perm[j] = p.perm[j - 1];
      p.perm[j - 1] = ti_;
    }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    const int max_smem = static_cast<int>(wfk::metrics_geom(1, 64 * wfk::kMaxColWarps).smem);
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::metrics_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::metrics_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_lock.finished();
  }
  if (p.vec2)
    wfk::metrics_strip_kernel<true><<<dim3(g.ns, g.nseg, frames), 32 * g.nwc, g.smem, s>>>(p, static_cast<wfk::TileRec*>(workspace));
  else
    wfk::metrics_strip_kernel<false><<<dim3(g.ns, g.nseg, frames), 32 * g.nwc, g.smem, s>>>(p, static_cast<wfk::TileRec*>(workspace));
  int rc = wfk::launched("metrics_strip_kernel");
  if (rc != WFK_OK) return rc;
  const size_t tile_bytes = (static_cast<size_t>(g.tiles()) * frames * sizeof(wfk::TileRec) + 255) & ~static_cast<size_t>(255);
  wfk::FrameRec* frs = reinterpret_cast<wfk::FrameRec*>(static_cast<uint8_t*>(workspace) + tile_bytes);
  wfk::metrics_frame_kernel<<<frames, 128, 0, s>>>(static_cast<const wfk::TileRec*>(workspace), g.tiles(), h, w, frs);
  rc = wfk::launched("metrics_frame_kernel");
  if (rc != WFK_OK) return rc;
  wfk::ThrPerm perm;
  for (int i = 0; i < WFK_MAX_THRESHOLDS; ++i) perm.v[i] = p.perm[i];
  wfk::metrics_finalize_kernel<<<1, 1024, 0, s>>>(frs, frames, n_thresholds, perm, out);
  return wfk::launched("metrics_finalize_kernel");
}
