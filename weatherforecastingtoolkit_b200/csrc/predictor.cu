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

// ---------------------------------------------------------------------------------------------
// Tensor-core version for the Path-B shape family (C = 4 latent channels, hw % 4 == 0): the per-pixel
// [1 x K] x [K x N] products of 32 pixels form an M = 32 GEMM tile per warp, run as mma.sync m16n8k8 TF32
// with the 3xTF32 split (x = hi + lo, both TF32; hi*hi + hi*lo + lo*hi accumulates in fp32), which keeps
// fp32-grade accuracy (~1e-6 relative; a single TF32 pass would be ~5e-4). The scalar kernel above needs two
// shared-memory operands per FMA and ran at 0.4 % of the HBM roofline; here every global access is a 128-bit
// load / store of 4 consecutive pixels of one latent plane (a warp covers 4 planes x 128 contiguous bytes),
// the 10 KB weight matrix is split and laid out in mma B-fragment order once per CTA.
//   M-tile rows are PERMUTED: rows (g, g + 8) of m-tile 0 and of m-tile 1 are pixels 4g .. 4g + 3, so that
//   the A fragments (a0, a1 | a0', a1') and the C fragments of one lane are one float4 each.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kPredTcThreads = 128;   // 4 warps x 32 pixels

template <int KS, int NT>   // K <= 8*KS, N == 8*NT
__global__ void __launch_bounds__(kPredTcThreads, 3) predict_linear_tc_kernel(
    const float* __restrict__ lat, const float* __restrict__ weight, const float* __restrict__ bias, int b, int t_in,
    int t_out, int hw, float* __restrict__ pred, float* __restrict__ tgt, double* __restrict__ loss_sums) {
  constexpr int C = 4;
  __shared__ uint4 s_b[KS * NT * 32];   // per (k-step, n-tile, lane): (b0_hi, b1_hi, b0_lo, b1_lo)
  __shared__ float s_w[KS * 8 * NT * 8];  // raw weights, staged with coalesced loads first
  __shared__ float s_bias[NT * 8];
  __shared__ float s_red[kPredTcThreads / 32];
  const int K = t_in * C, N = t_out * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // weights: one coalesced pass global -> shared (all loads in flight at once), then the split / fragment layout
  // from shared memory. (Building the fragments straight from global memory cost 42 dependent L2 round trips per
  // thread: 17 us of a 50 us kernel.)
  {
    constexpr int kPer = (KS * 8 * NT * 8 + kPredTcThreads - 1) / kPredTcThreads;
    float wv[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = threadIdx.x + j * kPredTcThreads;
      wv[j] = (i < N * K) ? __ldg(weight + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = threadIdx.x + j * kPredTcThreads;
      if (i < KS * 8 * NT * 8) s_w[i] = wv[j];
    }
  }
  for (int i = threadIdx.x; i < NT * 8; i += blockDim.x) s_bias[i] = i < N ? __ldg(bias + i) : 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
    const int n = nt * 8 + (ln >> 2), k0 = ks * 8 + (ln & 3), k1 = k0 + 4;
    const float w0 = (n < N && k0 < K) ? s_w[n * K + k0] : 0.f;
    const float w1 = (n < N && k1 < K) ? s_w[n * K + k1] : 0.f;
    const uint32_t h0 = to_tf32(w0), h1 = to_tf32(w1);
    s_b[i] = make_uint4(h0, h1, to_tf32(w0 - __uint_as_float(h0)), to_tf32(w1 - __uint_as_float(h1)));
  }
  __syncthreads();

  const int64_t total = static_cast<int64_t>(b) * hw;
  const int64_t tiles = (total + kPredTcThreads - 1) / kPredTcThreads;
  float loss = 0.f;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t idx = tile * kPredTcThreads + warp * 32 + 4 * g;   // first of this lane's 4 pixels
    const bool valid = idx < total;                                  // hw % 4 == 0: a group never straddles sequences
    const int bi = valid ? static_cast<int>(idx / hw) : 0;
    const int p = valid ? static_cast<int>(idx - static_cast<int64_t>(bi) * hw) : 0;
    const float* lb = lat + static_cast<int64_t>(bi) * (t_in + t_out) * C * hw + p;
    // last input frame, channel t (A operand rows k = 8*ks + t and + 4 are both channel t: C == 4)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // plane offsets as 32-bit element counts (one sequence of latents is far below 2^31 elements)
    const unsigned uhw = static_cast<unsigned>(hw);
    const float4 last_a = valid ? __ldg(reinterpret_cast<const float4*>(lb + static_cast<unsigned>((t_in - 1) * C + t) * uhw)) : zero4;
    float4 xa[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const int k = ks * 8 + t + 4 * hlf;
        xa[ks][hlf] = (valid && k < K) ? __ldg(reinterpret_cast<const float4*>(lb + static_cast<unsigned>(k) * uhw)) : zero4;
      }
    }
    // every other global read of the tile is issued here as well (target frames, the last input frame of the output
    // channels): loaded where they are used, the 12 target loads of the epilogue each cost a full memory round trip
    // (long-scoreboard stalls were 7 of every 8 warp cycles)
    float4 last_o[2], tvv[NT][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
      last_o[e] = valid ? __ldg(reinterpret_cast<const float4*>(lb + static_cast<unsigned>((t_in - 1) * C + ((2 * t + e) & 3)) * uhw)) : zero4;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int o = nt * 8 + 2 * t + e;
        tvv[nt][e] = (valid && o < N) ? __ldg(reinterpret_cast<const float4*>(lb + static_cast<unsigned>(K + o) * uhw)) : zero4;
      }
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      // residual framing (train.py:104-106); rows beyond K stay zero
      float x[2][4];   // [half][pixel]
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const bool on = (ks * 8 + t + 4 * hlf) < K;
        x[hlf][0] = on ? xa[ks][hlf].x - last_a.x : 0.f;
        x[hlf][1] = on ? xa[ks][hlf].y - last_a.y : 0.f;
        x[hlf][2] = on ? xa[ks][hlf].z - last_a.z : 0.f;
        x[hlf][3] = on ? xa[ks][hlf].w - last_a.w : 0.f;
      }
      uint32_t ah[2][4], al[2][4];   // [m-tile][a0..a3] = (row g, row g+8) x (col t, col t+4) = pixels (2mt, 2mt+1) x halves
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float v = x[r >> 1][2 * mt + (r & 1)];
          ah[mt][r] = to_tf32(v);
          al[mt][r] = to_tf32(v - __uint_as_float(ah[mt][r]));
        }
      }
      // 3xTF32: the three passes of one accumulator depend on each other (~30 cycles of HMMA latency each), so the
      // passes are the OUTER loop: 12 independent accumulators sit between two MMAs of the same one
      uint4 bf[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bf[nt] = s_b[(ks * NT + nt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], al[mt], bf[nt].x, bf[nt].y);   // lo * hi
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], ah[mt], bf[nt].z, bf[nt].w);   // hi * lo
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], ah[mt], bf[nt].x, bf[nt].y);   // hi * hi
    }
    float* pb = pred + static_cast<int64_t>(bi) * N * hw + p;
    float* tb = tgt + static_cast<int64_t>(bi) * N * hw + p;
    if (valid) {
      // outputs o = 8*nt + 2t + e: channel (2t + e) % 4; C fragment (c0, c2 | c0', c2') = pixels 0..3 for e = 0, (c1, c3 | ..) for e = 1
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int o = nt * 8 + 2 * t + e;
          if (o < N) {
            const float bo = s_bias[o];
            const float4 tv = tvv[nt][e];
            const float y0 = acc[0][nt][e] + bo, y1 = acc[0][nt][2 + e] + bo, y2 = acc[1][nt][e] + bo, y3 = acc[1][nt][2 + e] + bo;
            const float4 l = last_o[e];
            const float r0 = tv.x - l.x, r1 = tv.y - l.y, r2 = tv.z - l.z, r3 = tv.w - l.w;
            const unsigned oi = static_cast<unsigned>(o) * uhw;
            *reinterpret_cast<float4*>(pb + oi) = make_float4(y0 + l.x, y1 + l.y, y2 + l.z, y3 + l.w);
            if (tgt != nullptr) *reinterpret_cast<float4*>(tb + oi) = make_float4(r0 + l.x, r1 + l.y, r2 + l.z, r3 + l.w);
            const float d0 = y0 - r0, d1 = y1 - r1, d2 = y2 - r2, d3 = y3 - r3;
            loss = fmaf(d0, d0, loss);
            loss = fmaf(d1, d1, loss);
            loss = fmaf(d2, d2, loss);
            loss = fmaf(d3, d3, loss);
          }
        }
      }
    }
  }
  if (loss_sums != nullptr) {
    for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) s_red[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < kPredTcThreads / 32; ++i) tot += s_red[i];
      atomicAdd(&loss_sums[0], tot);
      if (blockIdx.x == 0) atomicAdd(&loss_sums[1], static_cast<double>(total) * N);
    }
  }
}

}  // namespace wfk

extern "C" int wfk_predict_linear(const float* lat, const float* weight, const float* bias, int b, int t_in, int t_out,
                                  int c, int hw, float* pred, float* tgt, double* loss_sums, void* stream) {
  WFK_ENTER(stream, lat);
  WFK_REQUIRE(lat && weight && bias && pred, "null pointer");
  WFK_REQUIRE(b > 0 && t_in > 0 && t_out > 0 && c > 0 && hw > 0, "empty problem");
  const int K = t_in * c, N = t_out * c;
  const uintptr_t align = reinterpret_cast<uintptr_t>(lat) | reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt);
  if (c == 4 && K <= 56 && N == 48 && hw % 4 == 0 && (align & 15) == 0) {
    // Path-B shape (13 -> 12 frames, 4 channels): tensor-core kernel, persistent over 128-pixel tiles
    const int64_t tiles = (static_cast<int64_t>(b) * hw + wfk::kPredTcThreads - 1) / wfk::kPredTcThreads;
    const int64_t cap = static_cast<int64_t>(wfk::num_sms()) * 3;   // 3 CTAs per SM are resident (<= 168 registers)
    const unsigned blocks = static_cast<unsigned>(tiles < cap ? tiles : cap);
    wfk::predict_linear_tc_kernel<7, 6><<<blocks, wfk::kPredTcThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        lat, weight, bias, b, t_in, t_out, hw, pred, tgt, loss_sums);
    return wfk::launched("predict_linear_tc_kernel");
  }
  const size_t smem = (static_cast<size_t>(N) * K + N + static_cast<size_t>(K) * wfk::kPredThreads) * sizeof(float);
  WFK_REQUIRE(smem <= 200 * 1024, "predictor too large for shared memory (K=%d N=%d)", K, N);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::predict_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_lock.finished();
  }
  const int64_t total = static_cast<int64_t>(b) * hw;
  const unsigned blocks = static_cast<unsigned>((total + wfk::kPredThreads - 1) / wfk::kPredThreads);
  wfk::predict_linear_kernel<<<blocks, wfk::kPredThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      lat, weight, bias, b, t_in, t_out, c, hw, pred, tgt, loss_sums);
  return wfk::launched("predict_linear_kernel");
}
