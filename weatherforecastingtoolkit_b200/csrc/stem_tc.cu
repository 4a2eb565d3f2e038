// 3x3 (pad 1, stride 1) "stem" convolutions with a tiny input-channel count on the tensor cores
// (legacy mma.sync m16n8k16, fp16 operands / fp32 accumulate like every other convolution of the model):
//   encoder.conv_in  1 -> 128 @ 384^2                          (reference pipeline/models/autoencoderkl/vae.py:24, 72)
//   post_quant_conv (1x1, 4 -> 4) + decoder.conv_in 4 -> 512   (autoencoder_kl.py:87; vae.py:103, 152)
// The contraction length is K = cin*9 (9 or, with the folded 1x1 and its bias as a constant-one plane, 45), so the
// im2col A fragment of 16 pixels is gathered straight from the fp32 NCHW input into registers (no staging), the
// [K x 128] weight fragments live in shared memory, and one warp produces 16 pixels x 128 channels per step.
// The CUDA-core kernels they replace (edge_convs.cu / in_conv.cu) needed 9*cin FMAs per output value and ran 5-13x
// above the HBM time of the layer. Also accumulates the GroupNorm (sum, sum of squares) of the fp32 result.
#include "act16.cuh"
#include "internal.h"

namespace wfk {

constexpr int kStemTcWarps = 4;
constexpr int kStemTcThreads = kStemTcWarps * 32;
constexpr int kStemTcN = 128;           // channels per block (16 n8-tiles)
constexpr int kStemTcPitch = kStemTcN + 8;

// in [n, cin, h, w] fp32; wt [KSTEPS*16][cout] fp16 with row k = ci*9 + tap (ci = cin is the constant-one plane when
// `ones_plane`), zero rows beyond K; out [n, h, w, cout] fp16; stats [n][cout/cpg][2] double.
template <int KSTEPS, bool BF16>
__global__ void __launch_bounds__(kStemTcThreads, KSTEPS == 1 ? 5 : 3) conv3x3_stem_tc_kernel(
    const float* __restrict__ in, int cin, int ones_plane, int h, int w, const uint16_t* __restrict__ wt,
    const float* __restrict__ bias, int cout, uint16_t* __restrict__ out, double* __restrict__ stats, int cpg,
    int tiles_per_warp) {
  __shared__ uint2 s_bf[KSTEPS][16][32];
  __shared__ __align__(16) uint16_t s_tile[kStemTcWarps][16][kStemTcPitch];
  __shared__ float s_bias[kStemTcN];
  __shared__ float s_stats[kStemTcWarps][kStemTcN / 4][2];
  const int n = blockIdx.y, chunk = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int hw = h * w;
  const int K = (cin + (ones_plane ? 1 : 0)) * 9;
  for (int i = threadIdx.x; i < KSTEPS * 16 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) & 15, ks = i >> 9;
    const int col = chunk * kStemTcN + nt * 8 + (ln >> 2);
    const int k0 = ks * 16 + (ln & 3) * 2;
    // two 16-bit weights per register (k, k + 1): pure bit packing, the same for fp16 and bf16
    const uint32_t p0 = wt[static_cast<int64_t>(k0) * cout + col] | (static_cast<uint32_t>(wt[static_cast<int64_t>(k0 + 1) * cout + col]) << 16);
    const uint32_t p1 = wt[static_cast<int64_t>(k0 + 8) * cout + col] | (static_cast<uint32_t>(wt[static_cast<int64_t>(k0 + 9) * cout + col]) << 16);
    s_bf[ks][nt][ln] = make_uint2(p0, p1);
  }
  for (int i = threadIdx.x; i < kStemTcN; i += blockDim.x) s_bias[i] = bias[chunk * kStemTcN + i];
  for (int i = threadIdx.x; i < kStemTcWarps * (kStemTcN / 4) * 2; i += blockDim.x) (&s_stats[0][0][0])[i] = 0.f;
  __syncthreads();
  // this lane's four k indices per k-step: 2t, 2t+1, 2t+8, 2t+9 -> (plane, dy, dx) packed; -1 = zero padding of K
  int kinfo[KSTEPS][4];
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ks * 16 + 2 * t + (j & 1) + ((j >> 1) << 3);
      if (k < K) {
        const int ci = k / 9, tap = k - 9 * ci;
        kinfo[ks][j] = (ci << 8) | ((tap / 3) << 4) | (tap % 3);
      } else {
        kinfo[ks][j] = -1;
      }
    }
  const float* inn = in + static_cast<int64_t>(n) * cin * hw;
  float ssum[16], ssq[16];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) ssum[nt] = ssq[nt] = 0.f;
  const int mtiles = (hw + 15) >> 4;
  const int mt0 = (blockIdx.x * kStemTcWarps + warp) * tiles_per_warp;
  for (int i = 0; i < tiles_per_warp; ++i) {
    const int mt = mt0 + i;
    if (mt >= mtiles) break;
    const int p_lo = mt * 16 + g, p_hi = p_lo + 8;
    const bool v_lo = p_lo < hw, v_hi = p_hi < hw;
    const int y_lo = p_lo / w, x_lo = p_lo - y_lo * w;
    const int y_hi = p_hi / w, x_hi = p_hi - y_hi * w;
    float acc[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      // A fragment: a0 = rows g, k (2t, 2t+1); a1 = rows g+8, same k; a2 = rows g, k + 8; a3 = rows g+8, k + 8
      float va[4][2];  // [j][row lo/hi]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ki = kinfo[ks][j];
        float lo = 0.f, hi = 0.f;
        if (ki >= 0) {
          const int ci = ki >> 8, dy = ((ki >> 4) & 15) - 1, dx = (ki & 15) - 1;
          const int yl = y_lo + dy, xl = x_lo + dx, yh = y_hi + dy, xh = x_hi + dx;
          const bool il = v_lo && yl >= 0 && yl < h && xl >= 0 && xl < w;
          const bool ih = v_hi && yh >= 0 && yh < h && xh >= 0 && xh < w;
          if (ci < cin) {
            if (il) lo = __ldg(inn + static_cast<int64_t>(ci) * hw + yl * w + xl);
            if (ih) hi = __ldg(inn + static_cast<int64_t>(ci) * hw + yh * w + xh);
          } else {  // constant-one plane: carries the folded 1x1 bias, present only where the tap is inside the image
            lo = il ? 1.f : 0.f;
            hi = ih ? 1.f : 0.f;
          }
        }
        va[j][0] = lo;
        va[j][1] = hi;
      }
      uint32_t a[4];
      {
        a[0] = A16<BF16>::pack(va[0][0], va[1][0]);
        a[1] = A16<BF16>::pack(va[0][1], va[1][1]);
        a[2] = A16<BF16>::pack(va[2][0], va[3][0]);
        a[3] = A16<BF16>::pack(va[2][1], va[3][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const uint2 b = s_bf[ks][nt][lane];
        mma_16816<BF16>(acc[nt], a, b.x, b.y);
      }
    }
    // epilogue: + bias, GroupNorm partial sums of the fp32 values, fp16 tile in this warp's shared memory
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float2 bb = *reinterpret_cast<const float2*>(&s_bias[nt * 8 + 2 * t]);
      const float c0 = acc[nt][0] + bb.x, c1 = acc[nt][1] + bb.y, c2 = acc[nt][2] + bb.x, c3 = acc[nt][3] + bb.y;
      if (v_lo) {
        ssum[nt] += c0 + c1;
        ssq[nt] = fmaf(c0, c0, fmaf(c1, c1, ssq[nt]));
      }
      if (v_hi) {
        ssum[nt] += c2 + c3;
        ssq[nt] = fmaf(c2, c2, fmaf(c3, c3, ssq[nt]));
      }
      *reinterpret_cast<uint32_t*>(&s_tile[warp][g][nt * 8 + 2 * t]) = A16<BF16>::pack(c0, c1);
      *reinterpret_cast<uint32_t*>(&s_tile[warp][g + 8][nt * 8 + 2 * t]) = A16<BF16>::pack(c2, c3);
    }
    __syncwarp();
    // coalesced NHWC store: 16 pixels x 256 B, 16 lanes per pixel
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 32 + lane;
      const int row = idx >> 4, ck = idx & 15;
      const int p = mt * 16 + row;
      if (p < hw)
        *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(n) * hw + p) * cout + chunk * kStemTcN + ck * 8) =
            *reinterpret_cast<const uint4*>(&s_tile[warp][row][ck * 8]);
    }
  }
  if (stats != nullptr) {
    // lanes that share a channel pair differ in g (lane bits 2..4); the pair's neighbour in a 4-channel group is
    // lane bit 0; 8- and 16-channel groups also fold lane bit 1
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      float s = ssum[nt], q = ssq[nt];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      q += __shfl_xor_sync(0xffffffffu, q, 1);
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (cpg >= 8) {
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
      }
      const bool writer = (cpg >= 8) ? (lane == 0) : (lane == 0 || lane == 2);
      if (writer) {  // this warp's private slots: plain read-modify-write in a fixed order
        const int gl = (nt * 8 + 2 * t) / cpg;
        s_stats[warp][gl][0] += s;
        s_stats[warp][gl][1] += q;
      }
      __syncwarp();
    }
    __syncthreads();
    const int groups_chunk = kStemTcN / cpg;
    if (threadIdx.x < 2 * groups_chunk) {
      const int gl = threadIdx.x >> 1, m = threadIdx.x & 1;
      float tot = 0.f;
#pragma unroll
      for (int wi = 0; wi < kStemTcWarps; ++wi) tot += s_stats[wi][gl][m];
      const int groups = cout / cpg;
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + chunk * groups_chunk + gl) * 2 + m], static_cast<double>(tot));
    }
  }
}

}  // namespace wfk

extern "C" int wfk_conv3x3_stem_tc(const float* in, int n, int cin, int h, int w, int ones_plane, const void* weight_h,
                                   const float* bias, int cout, void* out, double* stats, int cpg, int bf16, void* stream) {
  WFK_ENTER(stream, in);
  WFK_REQUIRE(in && weight_h && bias && out, "null pointer");
  WFK_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && cin >= 1, "bad shape");
  const int K = (cin + (ones_plane ? 1 : 0)) * 9;
  WFK_REQUIRE(K <= 48, "cin=%d (+%d) gives K=%d > 48: use the tcgen05 conv-GEMM", cin, ones_plane ? 1 : 0, K);
  WFK_REQUIRE(cout % wfk::kStemTcN == 0 && cout / wfk::kStemTcN <= 65535, "cout=%d must be a multiple of 128", cout);
  if (stats) WFK_REQUIRE(cpg == 4 || cpg == 8 || cpg == 16, "cpg=%d unsupported (4, 8, 16)", cpg);
  const int ksteps = (K + 15) / 16;
  const int mtiles = (h * w + 15) / 16;
  int tpw = 8;
  while (tpw > 1 && static_cast<int64_t>((mtiles + wfk::kStemTcWarps * tpw - 1) / (wfk::kStemTcWarps * tpw)) * n * (cout / wfk::kStemTcN) <
                        4 * static_cast<int64_t>(wfk::num_sms()))
    tpw >>= 1;
  dim3 grid((mtiles + wfk::kStemTcWarps * tpw - 1) / (wfk::kStemTcWarps * tpw), n, cout / wfk::kStemTcN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint16_t* wh = static_cast<const uint16_t*>(weight_h);
  uint16_t* oh = static_cast<uint16_t*>(out);
#define WFK_STEM_LAUNCH(KS, BF) \
  wfk::conv3x3_stem_tc_kernel<KS, BF><<<grid, wfk::kStemTcThreads, 0, s>>>(in, cin, ones_plane, h, w, wh, bias, cout, oh, stats, cpg, tpw)
  if (bf16) {
    if (ksteps == 1) WFK_STEM_LAUNCH(1, true);
    else if (ksteps == 2) WFK_STEM_LAUNCH(2, true);
    else WFK_STEM_LAUNCH(3, true);
  } else {
    if (ksteps == 1) WFK_STEM_LAUNCH(1, false);
    else if (ksteps == 2) WFK_STEM_LAUNCH(2, false);
    else WFK_STEM_LAUNCH(3, false);
  }
#undef WFK_STEM_LAUNCH
  return wfk::launched("conv3x3_stem_tc_kernel");
}
