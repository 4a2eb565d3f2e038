// decoder tail: GroupNorm + SiLU + conv3x3(C -> 1) fused in one pass over the raw fp16 stream.
// Replaces conv_norm_out / conv_act / conv_out of Decoder.forward (reference
// pipeline/models/autoencoderkl/vae.py:162-164) -- previously a GroupNorm pass (read + write of the
// 128-channel 384x384 tensor) followed by a gather-style direct convolution.
//
// out[p] = bias + sum_tap sum_c w[tap][c] * silu(a[c]*x[p+tap][c] + b[c])   (zero outside the image)
// is evaluated as d[q][tap] = sum_c w[tap][c] * act(x[q][c]) for every pixel q of a 16x16 tile plus a
// 1-pixel halo (each input pixel is read ONCE, 256 contiguous bytes, by one thread), followed by the
// 9-value gather out[p] = sum_tap d[p + tap][tap] through shared memory. HBM-bound: C*2 bytes per pixel.
#include <cuda_fp16.h>

#include <cstdlib>

#include "act16.cuh"
#include "internal.h"

namespace wfk {

constexpr int kOT = 16;                 // output tile edge
constexpr int kOH = kOT + 2;            // with halo
constexpr int kOutThreads = 352;        // >= kOH*kOH = 324

// ---------------------------------------------------------------------------------------------------------
// Tensor-core version (legacy mma.sync m16n8k16, 16-bit operands / fp32 accumulate like every other conv of the
// model): the per-pixel contraction d[q][tap] = sum_c w[tap][c] * act(x[q][c]) is a [pixels x C] x [C x 9] GEMM.
// Each warp takes 16 halo pixels at a time. The contraction does not care in which ORDER the channels are summed, so
// the K slots of the MMA are assigned to channels such that the 16 bytes a lane loads (8 consecutive channels of
// pixel g = lane/4, chunk t = lane%4 of every 32-channel block) ARE its A-fragment registers of two K steps:
//   K step 2j+u, slot 2t+i   <-> channel 32j + 8t + 4u + i        (a0: pixel g, a1: pixel g+8)
//   K step 2j+u, slot 2t+8+i <-> channel 32j + 8t + 4u + 2 + i    (a2: pixel g, a3: pixel g+8)
// with the weight fragments permuted to match. No staging tile, no ldmatrix: load -> GroupNorm + SiLU in registers
// -> MMA. The first tensor-core version staged the activated tile in shared memory (600 instructions per 16 pixels,
// SiLU as two fp32 MUFU ops per pair: 49 % MUFU-pipe, 55 % issue utilisation at 2.9x the HBM time of the pass); this
// one needs ~400 including addressing. In fp16 mode h + h*tanh(h) is evaluated on the packed pair (tanh.approx.f16x2 +
// one HFMA2; the result is rounded to fp16 for the MMA anyway, the affine part stays fp32) -- note that the packed tanh
// still issues two MUFU.TANH.F16 operations, so the MUFU pipe (one op per element, 52 % busy) remains the floor.
constexpr int kTcWarps = 7;                       // 21 m16-tiles of halo pixels per block = 3 rounds of 7 warps
constexpr int kTcThreads = kTcWarps * 32;
constexpr int kTcMTiles = (kOH * kOH + 15) / 16;  // 21
constexpr int kTcStrip = 4;                       // tiles along x per block

// silu(z) for a channel pair, z = 2 * (x * a + b) (a, b pre-halved), as a packed 16-bit pair
template <bool BF16>
__device__ __forceinline__ uint32_t silu_pair(uint32_t x2, float2 a, float2 b) {
  const float2 hv = __ffma2_rn(A16<BF16>::unpack(x2), a, b);
  if constexpr (BF16) {
    float2 tv;
    asm("tanh.approx.f32 %0, %1;" : "=f"(tv.x) : "f"(hv.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(tv.y) : "f"(hv.y));
    const float2 yv = __ffma2_rn(hv, tv, hv);
    return A16<true>::pack(yv.x, yv.y);
  } else {
    const uint32_t h2 = A16<false>::pack(hv.x, hv.y);
    uint32_t t2, y2;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t2) : "r"(h2));
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y2) : "r"(h2), "r"(t2));
    return y2;
  }
}

template <int KSTEPS, bool BF16>
__global__ void __launch_bounds__(kTcThreads, KSTEPS <= 8 ? 4 : 2) gn_silu_conv3x3_c1_tc_kernel(
    const uint16_t* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, int h, int w, int groups, float eps, const float* __restrict__ wt, float bias,
    float* __restrict__ out, int* __restrict__ nonfinite) {
  constexpr int C = KSTEPS * 16;
  constexpr int kBlocks32 = C / 32;      // 32-channel blocks = 16-byte loads per pixel and lane
  static_assert(KSTEPS % 2 == 0, "C must be a multiple of 32");
  extern __shared__ __align__(16) uint8_t s_raw[];
  float* s_a = reinterpret_cast<float*>(s_raw);        // [C]  (already halved: silu(z) = h + h*tanh(h), h = z/2)
  float* s_b = s_a + C;                                // [C]
  float* s_d = s_b + C;                                // [kTcMTiles*16][9]
  uint2* s_bf = reinterpret_cast<uint2*>(s_d + kTcMTiles * 16 * 9);       // [KSTEPS][2][32] weight fragments
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ int s_pixoff[kTcMTiles * 16];   // halo pixel q -> element offset of its channel 0 relative to the tile origin
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kOT;
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  if (threadIdx.x < groups) {
    const double cnt = static_cast<double>(cpg) * h * w;
    const int gi = threadIdx.x;
    const double sum = stats[(static_cast<int64_t>(n) * groups + gi) * 2 + 0];
    const double sq = stats[(static_cast<int64_t>(n) * groups + gi) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[gi] = static_cast<float>(mean);
    s_rstd[gi] = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const float a = s_rstd[ch / cpg] * gamma[ch];
    s_a[ch] = 0.5f * a;
    s_b[ch] = 0.5f * (beta[ch] - s_mean[ch / cpg] * a);
  }
  for (int q = threadIdx.x; q < kTcMTiles * 16; q += blockDim.x) {
    const int qq = min(q, kOH * kOH - 1);   // rows past the halo square re-read its last pixel (their d values are unused)
    s_pixoff[q] = ((qq / kOH - 1) * w + (qq % kOH - 1)) * C;
  }
  // weight fragments (B operand, "col" layout) staged once per block as [ks][nt][lane] uint2 with the K-slot -> channel
  // assignment above: lane = 4*n' + t', tap = nt*8 + n' (taps 9..15 are zero padding), ks = 2j + u:
  //   b0 = {W[tap][32j + 8t' + 4u], W[.. + 1]},  b1 = {W[.. + 2], W[.. + 3]}
  for (int i = threadIdx.x; i < KSTEPS * 2 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 1, ks = i >> 6;
    const int tap = nt * 8 + (ln >> 2);
    const int c0 = (ks >> 1) * 32 + (ln & 3) * 8 + (ks & 1) * 4;
    uint2 b = make_uint2(0u, 0u);
    if (tap < 9) {
      b.x = A16<BF16>::pack(__ldg(wt + tap * C + c0), __ldg(wt + tap * C + c0 + 1));
      b.y = A16<BF16>::pack(__ldg(wt + tap * C + c0 + 2), __ldg(wt + tap * C + c0 + 3));
    }
    s_bf[i] = b;
  }
  __syncthreads();
  // A block walks a strip of kTcStrip tiles along x (prologue paid once); each warp owns m-tiles warp, warp+7, warp+14
  // of every tile.
  const int tiles_x = (w + kOT - 1) / kOT;
  const int tx_first = blockIdx.x * kTcStrip;
  const int ntiles = min(kTcStrip, tiles_x - tx_first);
  constexpr int kPerTile = kTcMTiles / kTcWarps;   // 3
  static_assert(kTcMTiles % kTcWarps == 0, "m-tiles must split evenly over the warps");
  const int n_items = ntiles * kPerTile;
  // One m-tile of work for this lane: the 16-byte chunk t of every 32-channel block of halo pixels g (lo) and g + 8 (hi)
  struct Item {
    const uint16_t* p_lo;
    const uint16_t* p_hi;
    bool ok_lo, ok_hi;
    int mt;
  };
  auto locate = [&](const int item) {
    Item it;
    const int ts = item / kPerTile;
    it.mt = warp + kTcWarps * (item - ts * kPerTile);
    const int x0 = (tx_first + ts) * kOT;
    const int q_lo = it.mt * 16 + g, q_hi = q_lo + 8;
    if (y0 >= 1 && y0 + kOT < h && x0 >= 1 && x0 + kOT < w) {
      // interior tile: every halo pixel is inside the image; the address is the tile origin plus a table entry
      const uint16_t* origin = x + ((static_cast<int64_t>(n) * h + y0) * w + x0) * C + t * 8;
      it.p_lo = origin + s_pixoff[q_lo];
      it.p_hi = origin + s_pixoff[q_hi];
      it.ok_lo = it.ok_hi = true;
    } else {
      auto addr = [&](const int q, bool& ok) {
        const int qy = q / kOH, qx = q - qy * kOH;
        const int y = y0 + qy - 1, xx = x0 + qx - 1;
        ok = q < kOH * kOH && y >= 0 && y < h && xx >= 0 && xx < w;
        const int yc = min(max(y, 0), h - 1), xc = min(max(xx, 0), w - 1);   // clamped: the load itself is unconditional
        return x + ((static_cast<int64_t>(n) * h + yc) * w + xc) * C + t * 8;
      };
      it.p_lo = addr(q_lo, it.ok_lo);
      it.p_hi = addr(q_hi, it.ok_hi);
    }
    return it;
  };
  // Rolling prefetch: as soon as block j of the current m-tile has been consumed, the same registers receive block j of
  // the NEXT m-tile (possibly of the next tile), so every warp keeps 8 loads in flight through its MMAs, the gather and
  // the block barriers without a second register buffer. (Loading an m-tile and then processing it left the pass
  // latency-bound at 40 % of the HBM rate even after the instruction count had been cut by half.)
  auto process = [&](const Item& cur, const Item& nxt, const bool has_next, uint4 (&lo)[kBlocks32], uint4 (&hi)[kBlocks32],
                     float* sd) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int j = 0; j < kBlocks32; ++j) {
      // (scale, shift) of this lane's 8 channels of block j: the four lanes of a pixel read four different rows, lanes
      // with equal t the same one (broadcast)
      const float4 a0 = *reinterpret_cast<const float4*>(s_a + j * 32 + t * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(s_a + j * 32 + t * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_b + j * 32 + t * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(s_b + j * 32 + t * 8 + 4);
      uint4 vl = make_uint4(0u, 0u, 0u, 0u), vh = make_uint4(0u, 0u, 0u, 0u);   // out-of-image pixels: zero AFTER the activation
      if (cur.ok_lo) {
        vl.x = silu_pair<BF16>(lo[j].x, make_float2(a0.x, a0.y), make_float2(b0.x, b0.y));
        vl.y = silu_pair<BF16>(lo[j].y, make_float2(a0.z, a0.w), make_float2(b0.z, b0.w));
        vl.z = silu_pair<BF16>(lo[j].z, make_float2(a1.x, a1.y), make_float2(b1.x, b1.y));
        vl.w = silu_pair<BF16>(lo[j].w, make_float2(a1.z, a1.w), make_float2(b1.z, b1.w));
      }
      if (cur.ok_hi) {
        vh.x = silu_pair<BF16>(hi[j].x, make_float2(a0.x, a0.y), make_float2(b0.x, b0.y));
        vh.y = silu_pair<BF16>(hi[j].y, make_float2(a0.z, a0.w), make_float2(b0.z, b0.w));
        vh.z = silu_pair<BF16>(hi[j].z, make_float2(a1.x, a1.y), make_float2(b1.x, b1.y));
        vh.w = silu_pair<BF16>(hi[j].w, make_float2(a1.z, a1.w), make_float2(b1.z, b1.w));
      }
      if (has_next) {
        lo[j] = __ldg(reinterpret_cast<const uint4*>(nxt.p_lo + j * 32));
        hi[j] = __ldg(reinterpret_cast<const uint4*>(nxt.p_hi + j * 32));
      }
      {
        const uint2 bf0 = s_bf[((2 * j) * 2 + 0) * 32 + lane], bf1 = s_bf[((2 * j) * 2 + 1) * 32 + lane];
        const uint32_t a[4] = {vl.x, vh.x, vl.y, vh.y};
        mma_16816<BF16>(acc[0], a, bf0.x, bf0.y);
        mma_16816<BF16>(acc[1], a, bf1.x, bf1.y);
      }
      {
        const uint2 bf0 = s_bf[((2 * j + 1) * 2 + 0) * 32 + lane], bf1 = s_bf[((2 * j + 1) * 2 + 1) * 32 + lane];
        const uint32_t a[4] = {vl.z, vh.z, vl.w, vh.w};
        mma_16816<BF16>(acc[0], a, bf0.x, bf0.y);
        mma_16816<BF16>(acc[1], a, bf1.x, bf1.y);
      }
    }
    // C fragment: rows lane/4 and lane/4 + 8, taps nt*8 + (lane%4)*2 + {0, 1}
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tap = nt * 8 + t * 2 + (j & 1);
        const int row = g + ((j >> 1) << 3);
        if (tap < 9) sd[(cur.mt * 16 + row) * 9 + tap] = acc[nt][j];
      }
  };
  // after a tile's m-tiles: block barrier, the 9-tap gather from s_d, and a second barrier before the next tile's
  // m-tiles overwrite it (a second s_d buffer cost a resident block per SM and bought nothing)
  auto finish_tile = [&](const int ts, const float* sd) {
    __syncthreads();
    const int x0 = (tx_first + ts) * kOT;
    for (int p = threadIdx.x; p < kOT * kOT; p += blockDim.x) {
      const int py = p / kOT, pxx = p - py * kOT;
      const int y = y0 + py, xx = x0 + pxx;
      if (y < h && xx < w) {
        float acc = bias;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) acc += sd[((py + r) * kOH + (pxx + s)) * 9 + r * 3 + s];
        if (!isfinite(acc)) *nonfinite = 8;   // decoded frame: the last tensor of decode
        out[(static_cast<int64_t>(n) * h + y) * w + xx] = acc;
      }
    }
    __syncthreads();
  };
  // Measured on B200 inside the power-capped step (37 frames of 384^2): four resident blocks that load an m-tile and
  // then process it: 569 us; three blocks (80 registers) with the rolling prefetch: 626 us -- the pass is bound by the
  // MUFU pipe (one tanh per element, 1.27x of them because of the halo: 52 % busy next to 52 % issue utilisation), not
  // by load latency, so occupancy wins. kRollingPrefetch stays as the measured alternative.
  constexpr bool kRollingPrefetch = false;
  uint4 lo[kBlocks32], hi[kBlocks32];
  Item cur = locate(0);
  auto load_item = [&](const Item& it) {
#pragma unroll
    for (int j = 0; j < kBlocks32; ++j) {
      lo[j] = __ldg(reinterpret_cast<const uint4*>(it.p_lo + j * 32));
      hi[j] = __ldg(reinterpret_cast<const uint4*>(it.p_hi + j * 32));
    }
  };
  if (kRollingPrefetch) load_item(cur);
  for (int item = 0; item < n_items; ++item) {
    const int ts = item / kPerTile;
    const bool has_next = item + 1 < n_items;
    if (!kRollingPrefetch) load_item(cur);
    const Item nxt = has_next ? locate(item + 1) : cur;
    process(cur, nxt, kRollingPrefetch && has_next, lo, hi, s_d);
    if (item % kPerTile == kPerTile - 1) finish_tile(ts, s_d);
    cur = nxt;
  }
}

template <int KSTEPS, bool BF16>
int launch_tail_tc(const uint16_t* x, const double* stats, const float* gamma, const float* beta, int n, int h, int w,
                   int groups, float eps, const float* weight, float bias, float* out, cudaStream_t s) {
  constexpr int C = KSTEPS * 16;
  const size_t smem = (2 * C + kTcMTiles * 16 * 9) * sizeof(float) + static_cast<size_t>(KSTEPS) * 2 * 32 * 8;
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    cudaError_t e = cudaFuncSetAttribute(gn_silu_conv3x3_c1_tc_kernel<KSTEPS, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(WFK_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_lock.finished();
  }
  const int tiles_x = (w + kOT - 1) / kOT;
  dim3 grid((tiles_x + kTcStrip - 1) / kTcStrip, (h + kOT - 1) / kOT, n);
  gn_silu_conv3x3_c1_tc_kernel<KSTEPS, BF16><<<grid, kTcThreads, smem, s>>>(x, stats, gamma, beta, h, w, groups, eps, weight,
                                                                          bias, out, nonfinite_flag());
  return launched("gn_silu_conv3x3_c1_tc_kernel");
}


__global__ void __launch_bounds__(kOutThreads) gn_silu_conv3x3_c1_kernel(
    const __half* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, int h, int w, int c, int groups, float eps, const float* __restrict__ wt,
    float bias, float* __restrict__ out, int* __restrict__ nonfinite) {
  extern __shared__ float s_mem[];
  float* s_a = s_mem;                // [c]
  float* s_b = s_a + c;              // [c]
  float* s_w = s_b + c;              // [9][c]
  float* s_d = s_w + 9 * c;          // [kOH*kOH][9]
  float* s_mean = s_d + kOH * kOH * 9;
  float* s_rstd = s_mean + groups;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kOT, y0 = blockIdx.y * kOT;
  const int cpg = c / groups;
  if (threadIdx.x < groups) {
    const double cnt = static_cast<double>(cpg) * h * w;
    const int g = threadIdx.x;
    const double sum = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 0];
    const double sq = stats[(static_cast<int64_t>(n) * groups + g) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
  }
  for (int i = threadIdx.x; i < 9 * c; i += blockDim.x) s_w[i] = wt[i];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float a = s_rstd[ch / cpg] * gamma[ch];
    s_a[ch] = a;
    s_b[ch] = beta[ch] - s_mean[ch / cpg] * a;
  }
  __syncthreads();
  if (threadIdx.x < kOH * kOH) {
    const int qy = threadIdx.x / kOH, qx = threadIdx.x - qy * kOH;
    const int y = y0 + qy - 1, xx = x0 + qx - 1;
    float d[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) d[t] = 0.f;
    if (y >= 0 && y < h && xx >= 0 && xx < w) {
      const uint4* px = reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * h + y) * w + xx) * c);
      for (int v = 0; v < (c >> 3); ++v) {
        const uint4 u = __ldg(px + v);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        float act[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          const float z0 = fmaf(f.x, s_a[8 * v + 2 * e], s_b[8 * v + 2 * e]);
          const float z1 = fmaf(f.y, s_a[8 * v + 2 * e + 1], s_b[8 * v + 2 * e + 1]);
          act[2 * e] = __fdividef(z0, 1.f + __expf(-z0));
          act[2 * e + 1] = __fdividef(z1, 1.f + __expf(-z1));
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * c + 8 * v);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + t * c + 8 * v + 4);
          d[t] = fmaf(act[0], w0.x, d[t]);
          d[t] = fmaf(act[1], w0.y, d[t]);
          d[t] = fmaf(act[2], w0.z, d[t]);
          d[t] = fmaf(act[3], w0.w, d[t]);
          d[t] = fmaf(act[4], w1.x, d[t]);
          d[t] = fmaf(act[5], w1.y, d[t]);
          d[t] = fmaf(act[6], w1.z, d[t]);
          d[t] = fmaf(act[7], w1.w, d[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_d[threadIdx.x * 9 + t] = d[t];
  }
  __syncthreads();
  if (threadIdx.x < kOT * kOT) {
    const int py = threadIdx.x / kOT, pxx = threadIdx.x - py * kOT;
    const int y = y0 + py, xx = x0 + pxx;
    if (y < h && xx < w) {
      float acc = bias;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) acc += s_d[((py + r) * kOH + (pxx + s)) * 9 + r * 3 + s];
      out[(static_cast<int64_t>(n) * h + y) * w + xx] = acc;
    }
  }
}

}  // namespace wfk

extern "C" int wfk_gn_silu_conv3x3_c1(const void* x, const double* stats, const float* gamma, const float* beta, int n,
                                      int h, int w, int c, int groups, float eps, const float* weight, float bias,
                                      float* out, int bf16, void* stream) {
  WFK_ENTER(stream, x);
  WFK_REQUIRE(x && stats && gamma && beta && weight && out, "null pointer");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0, "bad shape");
  WFK_REQUIRE(c % 8 == 0 && c > 0 && c <= 1024 && c % groups == 0 && groups <= 256, "unsupported c=%d groups=%d", c, groups);
  if (groups <= 64 && !std::getenv("WFK_TAIL_CUDA_CORES")) {
    const uint16_t* xh = static_cast<const uint16_t*>(x);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bf16) {
      if (c == 128) return wfk::launch_tail_tc<8, true>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
      if (c == 64) return wfk::launch_tail_tc<4, true>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
      if (c == 256) return wfk::launch_tail_tc<16, true>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
    } else {
      if (c == 128) return wfk::launch_tail_tc<8, false>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
      if (c == 64) return wfk::launch_tail_tc<4, false>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
      if (c == 256) return wfk::launch_tail_tc<16, false>(xh, stats, gamma, beta, n, h, w, groups, eps, weight, bias, out, st);
    }
  }
  WFK_REQUIRE(!bf16, "bf16 activations: the CUDA-core tail kernel is fp16 only (c must be 64, 128 or 256)");
  const size_t smem = (static_cast<size_t>(c) * 11 + wfk::kOH * wfk::kOH * 9 + 2 * groups) * sizeof(float);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::gn_silu_conv3x3_c1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_lock.finished();
  }
  WFK_REQUIRE(smem <= 100 * 1024, "channel count too large for shared memory");
  dim3 grid((w + wfk::kOT - 1) / wfk::kOT, (h + wfk::kOT - 1) / wfk::kOT, n);
  wfk::gn_silu_conv3x3_c1_kernel<<<grid, wfk::kOutThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), stats, gamma, beta, h, w, c, groups, eps, weight, bias, out, wfk::nonfinite_flag());
  return wfk::launched("gn_silu_conv3x3_c1_kernel");
}
