// 3x3 (pad 1, stride 1) "stem" convolutions with a tiny input-channel count on the tensor cores
// (legacy mma.sync m16n8k16, fp16 operands / fp32 accumulate like every other convolution of the model):
//   encoder.conv_in  1 -> 128 @ 384^2                          (reference pipeline/models/autoencoderkl/vae.py:24, 72)
//   post_quant_conv (1x1, 4 -> 4) + decoder.conv_in 4 -> 512   (autoencoder_kl.py:87; vae.py:103, 152)
// The contraction length is K = cin*9 (9 or, with the folded 1x1 and its bias as a constant-one plane, 45), so the
// im2col A fragment of 16 pixels is gathered straight from the fp32 NCHW input into registers (no staging), the
// [K x 128] weight fragments live in shared memory, and one warp produces 16 pixels x 128 channels per step.
// The CUDA-core kernels they replace (edge_convs.cu / in_conv.cu) needed 9*cin FMAs per output value and ran 5-13x
// above the HBM time of the layer. Also accumulates the GroupNorm (sum, sum of squares) of the fp32 result.
#include <cstdlib>

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

// Single-plane variant (encoder.conv_in: 1 -> cout at 384^2, 37.7 MB of output per frame), for widths that are a
// multiple of 16 so that a 16-pixel M tile is a run of one image row. The generic kernel above spends ~900 instructions
// per tile (per-lane im2col addressing with bounds checks, scalar GroupNorm sums, a staged store) and runs at 3x the
// HBM time of the layer; this one needs ~200:
//  * the 3 x 18 input patch of a tile is loaded by 18 lanes (coalesced, one predicated load per row), parked as 16-bit
//    values in a double-buffered shared-memory patch, and every lane reads its A-fragment taps from fixed offsets;
//  * the bias rides through the tensor core: rows k = 9, 10, 11 of the weight fragment hold its hi / lo / lo2 16-bit
//    split (>= 24 significant bits) against constant-one A columns;
//  * MMA column n of n-tile nt maps to channel (nt>>2)*32 + (n>>1)*8 + (nt&3)*2 + (n&1): a lane's accumulators of four
//    consecutive n-tiles are 8 CONTIGUOUS channels, i.e. one 128-bit store per pixel row straight from registers (the
//    four lanes of a pixel cover 64 contiguous bytes: full 32-byte sectors, no staging tile);
//  * GroupNorm partial sums with packed fp32x2 adds / FMAs.
template <bool BF16>
__global__ void __launch_bounds__(kStemTcThreads, 4) conv3x3_stem1_tc_kernel(
    const float* __restrict__ in, int h, int w, const uint16_t* __restrict__ wt, const float* __restrict__ bias, int cout,
    uint16_t* __restrict__ out, double* __restrict__ stats, int cpg, int tiles_per_warp) {
  constexpr int kPatchPitch = 24;
  __shared__ uint2 s_bf[16][32];
  __shared__ uint16_t s_patch[kStemTcWarps][2][3 * kPatchPitch];
  __shared__ float s_stats[kStemTcWarps][kStemTcN / 4][2];
  const int n = blockIdx.y, chunk = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int hw = h * w;
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int ln = i & 31, nt = i >> 5;
    const int nn = ln >> 2;
    const int col = chunk * kStemTcN + (nt >> 2) * 32 + (nn >> 1) * 8 + (nt & 3) * 2 + (nn & 1);
    const int k0 = (ln & 3) * 2;
    auto wrow = [&](int k) -> uint32_t {
      if (k < 9) return wt[static_cast<int64_t>(k) * cout + col];
      if (k > 11) return 0u;
      // bias = hi + lo + lo2, each exactly representable in the 16-bit operand format
      float r = bias[col];
      uint16_t part = A16<BF16>::pack1(r);
      for (int j = 9; j < k; ++j) {
        r -= A16<BF16>::unpack1(part);
        part = A16<BF16>::pack1(r);
      }
      return part;
    };
    s_bf[nt][ln] = make_uint2(wrow(k0) | (wrow(k0 + 1) << 16), wrow(k0 + 8) | (wrow(k0 + 9) << 16));
  }
  for (int i = threadIdx.x; i < kStemTcWarps * (kStemTcN / 4) * 2; i += blockDim.x) (&s_stats[0][0][0])[i] = 0.f;
  __syncthreads();
  // fixed patch offsets of this lane's taps: k = 2t, 2t+1 (a0 / a1) and k = 8 (a2 / a3 of t == 0); patch column 0 is x0-1
  const int o0 = ((2 * t) / 3) * kPatchPitch + (2 * t) % 3 + g;
  const int o1 = ((2 * t + 1) / 3) * kPatchPitch + (2 * t + 1) % 3 + g;
  const int o8 = 2 * kPatchPitch + 2 + g;
  const uint32_t one16 = A16<BF16>::pack1(1.f);
  // a2 / a3 upper halves and constants: t == 0: {tap 8, 1}; t == 1: {1, 1}; t >= 2: 0
  const uint32_t a23_const = (t == 0) ? (one16 << 16) : (t == 1 ? (one16 | (one16 << 16)) : 0u);
  const float* inn = in + static_cast<int64_t>(n) * hw;
  float2 gs[8], gq[8];   // slot = (nt >> 2) * 2 + ((nt & 3) >> 1): 4 contiguous channels of this lane
#pragma unroll
  for (int i = 0; i < 8; ++i) gs[i] = gq[i] = make_float2(0.f, 0.f);
  const int mtiles = hw >> 4;
  const int mt0 = (blockIdx.x * kStemTcWarps + warp) * tiles_per_warp;
  const int mt_end = min(mt0 + tiles_per_warp, mtiles);
  int y = (mt0 * 16) / w, x0 = mt0 * 16 - y * w;
  auto load_patch = [&](int yy, int xx, float (&v)[3]) {
    const int col = xx - 1 + lane;
    const bool cok = lane < 18 && col >= 0 && col < w;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int row = yy - 1 + r;
      v[r] = (cok && row >= 0 && row < h) ? __ldg(inn + row * w + col) : 0.f;
    }
  };
  float v[3] = {0.f, 0.f, 0.f};
  if (mt0 < mt_end) load_patch(y, x0, v);
  uint16_t* outp = out + (static_cast<int64_t>(n) * hw + mt0 * 16 + g) * cout + chunk * kStemTcN + t * 8;
  const int64_t row8 = static_cast<int64_t>(8) * cout;
  for (int mt = mt0; mt < mt_end; ++mt) {
    uint16_t* pb = s_patch[warp][mt & 1];
    if (lane < 18) {
#pragma unroll
      for (int r = 0; r < 3; ++r) pb[r * kPatchPitch + lane] = A16<BF16>::pack1(v[r]);
    }
    // next tile's patch: in flight during this tile's MMAs and stores
    x0 += 16;
    if (x0 == w) {
      x0 = 0;
      ++y;
    }
    if (mt + 1 < mt_end) load_patch(y, x0, v);
    __syncwarp();
    uint32_t a[4];
    a[0] = pb[o0] | (static_cast<uint32_t>(pb[o1]) << 16);
    a[1] = pb[o0 + 8] | (static_cast<uint32_t>(pb[o1 + 8]) << 16);
    a[2] = a23_const | (t == 0 ? static_cast<uint32_t>(pb[o8]) : 0u);
    a[3] = a23_const | (t == 0 ? static_cast<uint32_t>(pb[o8 + 8]) : 0u);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const uint2 b = s_bf[q * 4 + j][lane];
        mma_16816<BF16>(acc[j], a, b.x, b.y);
      }
      uint4 lo, hi;
      uint32_t* lo32 = reinterpret_cast<uint32_t*>(&lo);
      uint32_t* hi32 = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 c01 = make_float2(acc[j][0], acc[j][1]), c23 = make_float2(acc[j][2], acc[j][3]);
        const int slot = q * 2 + (j >> 1);
        gs[slot] = __fadd2_rn(gs[slot], __fadd2_rn(c01, c23));
        gq[slot] = __ffma2_rn(c01, c01, gq[slot]);
        gq[slot] = __ffma2_rn(c23, c23, gq[slot]);
        lo32[j] = A16<BF16>::pack(c01.x, c01.y);
        hi32[j] = A16<BF16>::pack(c23.x, c23.y);
      }
      *reinterpret_cast<uint4*>(outp + q * 32) = lo;
      *reinterpret_cast<uint4*>(outp + row8 + q * 32) = hi;
    }
    outp += static_cast<int64_t>(16) * cout;
  }
  if (stats != nullptr) {
    // lanes that differ in g hold other pixels of the same channels: fold lane bits 2..4; then the lane's 8 slots of
    // 4 channels each are combined according to the group size (cpg 8: two slots; cpg 16: also lanes t, t^1)
#pragma unroll
    for (int slot = 0; slot < 8; ++slot) {
      float s = gs[slot].x + gs[slot].y, q = gq[slot].x + gq[slot].y;
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (cpg == 16) {
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
      }
      // channel of this slot inside the chunk: (slot >> 1) * 32 + t * 8 + (slot & 1) * 4
      const int gl = ((slot >> 1) * 32 + t * 8 + (slot & 1) * 4) / cpg;
      const bool writer = g == 0 && (cpg != 16 || (t & 1) == 0);
      if (writer) {  // this warp's private slots; for cpg >= 8 the two slots of a group are added one after the other
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
  // ONE wave of blocks: a block's prologue (weight-fragment table built from 16-bit global loads) costs as much as
  // ~10 tiles of work, so with a few tiles per warp it dominated both kernels (measured: encoder.conv_in 449 us for 37
  // frames with 8 tiles per warp, 336 us with one wave; the tile loop itself is ~1/3 of that).
  static const bool lean_enabled = !(std::getenv("WFK_STEM_LEAN") && std::getenv("WFK_STEM_LEAN")[0] == '0');   // A/B switch
  const bool lean = lean_enabled && cin == 1 && !ones_plane && w % 16 == 0 && static_cast<int64_t>(h) * w < (1 << 30);
  const int resident = lean ? 4 : (ksteps == 1 ? 5 : 3);   // blocks per SM (the kernels' __launch_bounds__)
  const int64_t slots = resident * static_cast<int64_t>(wfk::num_sms());
  const int64_t per_frame = slots / (static_cast<int64_t>(n) * (cout / wfk::kStemTcN));
  const int bpf = static_cast<int>(per_frame < 1 ? 1 : per_frame);   // blocks per (frame, 128-channel chunk)
  const int tpw = (mtiles + wfk::kStemTcWarps * bpf - 1) / (wfk::kStemTcWarps * bpf);
  dim3 grid((mtiles + wfk::kStemTcWarps * tpw - 1) / (wfk::kStemTcWarps * tpw), n, cout / wfk::kStemTcN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint16_t* wh = static_cast<const uint16_t*>(weight_h);
  uint16_t* oh = static_cast<uint16_t*>(out);
  if (lean) {   // single input plane, 16-pixel tiles inside one image row
    if (bf16)
      wfk::conv3x3_stem1_tc_kernel<true><<<grid, wfk::kStemTcThreads, 0, s>>>(in, h, w, wh, bias, cout, oh, stats, cpg, tpw);
    else
      wfk::conv3x3_stem1_tc_kernel<false><<<grid, wfk::kStemTcThreads, 0, s>>>(in, h, w, wh, bias, cout, oh, stats, cpg, tpw);
    return wfk::launched("conv3x3_stem1_tc_kernel");
  }
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
