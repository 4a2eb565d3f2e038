// ConvAttnModel latent compressor: the per-frame network of
// experiments/v1_experiments/pretrained_ae_convattn_ae_sevir/train.py:58-170 in ONE kernel, one CTA per latent frame,
// every activation resident in shared memory (three 144 x 128 fp32 token buffers, 223 KB):
//   encoder_cnn: Conv3x3 s2 p1 cin->64, GroupNorm(8), GELU, Conv3x3 s2 p1 64->128, GroupNorm(8), GELU   (48^2 -> 12^2)
//   tokens [144, 128] + encoder_pos_embedding; L x pre-norm TransformerEncoderLayer(128, 8 heads, ff 512, GELU)
//   attention_pool (one learned query over the 144 tokens), encoder_head LayerNorm + Linear(128 -> latent)  -> z
//   decoder_head Linear(latent -> 128) = the single memory token; queries = decoder_queries + decoder_pos_embedding;
//   L x pre-norm TransformerDecoderLayer; ConvT4x4 s2 p1 128->64, GroupNorm(8), GELU, ConvT4x4 s2 p1 64->cin (12^2 -> 48^2)
// plus the HuberLoss(pred, input) partial sums of the experiment's validation_step (train.py:172, 206-209).
// The reference issues a few hundred library kernels per call on a [B, 144, 128] problem (far below one wave of a
// B200); here the model is a single launch, the frame goes in and out of HBM once and the weights (6.5 MB) stay in L2.
// All arithmetic is fp32 on the CUDA cores: ~0.6 GFLOP per frame.
//
// Cross-attention over ONE memory token: softmax over a single key is 1, so the block's output is
// out_proj(v_proj(memory)) for every query -- a per-frame vector, independent of norm2(x) (train.py:150-156).
#include "internal.h"

namespace wfk {

constexpr int kCaThreads = 512;
constexpr int kCaWarps = kCaThreads / 32;
constexpr int kCaD = 128;   // transformer_embed_dim
constexpr int kCaT = 144;   // 12 x 12 tokens
constexpr int kCaDh = 16;   // 8 heads
constexpr int kCaFf = 512;  // dim_feedforward = 4 * embed
constexpr int kCaLd = 132;  // token row stride (floats): rows stay 16-byte aligned and 8 consecutive rows hit 8 distinct
                            // bank quads, so per-row float4 reads by different lanes are conflict-free
constexpr int kCaBuf = kCaT * kCaLd;  // 19,008 floats per token buffer
constexpr int kCaRowsPerThread = 9;   // 16 row groups x 9 rows = 144
constexpr int kCaMaxLayers = 8;
constexpr int kCaMaxLatent = 512;

// Matrices consumed by ca_gemm are stored TRANSPOSED ([in, out], packed once on the host) so that a warp's weight loads
// at one k are 512 contiguous bytes; matrices consumed by ca_linear keep the PyTorch [out, in] layout.
struct CaAttnW {
  const float *in_w, *in_b, *out_w, *out_b;  // in_proj (q | k | v along the output dimension), out_proj
};
struct CaEncLayerW {
  CaAttnW sa;
  const float *l1_w, *l1_b, *l2_w, *l2_b;  // linear1^T [128, 512], linear2^T [512, 128]
  const float *n1_w, *n1_b, *n2_w, *n2_b;
};
struct CaDecLayerW {
  CaAttnW sa, ca;
  const float *l1_w, *l1_b, *l2_w, *l2_b;
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;
};
struct ConvAttnWeights {
  const float *c0_w, *c0_b, *g0_w, *g0_b, *c1_w, *c1_b, *g1_w, *g1_b;
  const float* enc_pos;  // [144, 128]
  CaEncLayerW enc[kCaMaxLayers];
  const float* pool_q;  // [128]
  CaAttnW pool;            // [out, in] (the query / output projections are matrix-vector products)
  const float* pool_in_wt; // in_proj^T [128, 384] for the key / value GEMMs
  const float *eh_ln_w, *eh_ln_b, *eh_w, *eh_b;  // encoder_head: LayerNorm(128), Linear [latent, 128]
  const float *dh_w, *dh_b;                      // decoder_head [128, latent]
  const float *dec_q, *dec_pos;                  // [144, 128] each
  CaDecLayerW dec[kCaMaxLayers];
  const float *t0_w, *t0_b, *g2_w, *g2_b, *t1_w, *t1_b;  // ConvT [128, 64, 4, 4], GN(8, 64), ConvT [64, cin, 4, 4]
};

struct CaShared {
  float* buf[3];   // token buffers X, A, B (contiguous: the CNN stages treat them as one arena)
  float* mu;       // [144] LayerNorm row means
  float* rs;       // [144] LayerNorm row 1/std
  float* vec[3];   // [128] scratch vectors
  float* gstat;    // [16] GroupNorm (mean, rstd) x 8 groups
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// nn.LayerNorm(128, eps 1e-5) statistics of every token row: one warp per row.
__device__ void ca_row_stats(const float* x, float* mu, float* rs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kCaT; r += kCaWarps) {
    const float4 v = *reinterpret_cast<const float4*>(x + r * kCaLd + lane * 4);
    const float m = warp_sum(v.x + v.y + v.z + v.w) * (1.f / kCaD);
    const float a = v.x - m, b = v.y - m, c = v.z - m, d = v.w - m;
    const float q = warp_sum(a * a + b * b + c * c + d * d) * (1.f / kCaD);
    if (lane == 0) {
      mu[r] = m;
      rs[r] = rsqrtf(q + 1e-5f);
    }
  }
  __syncthreads();
}

enum { CA_STORE = 0, CA_STORE_GELU = 1, CA_ACC = 2 };

// out[144, n] (op)= f(in[144, 128]) W^T + bias with W^T given k-major: wt[k * ldw + col(c)] is the weight of input k for
// output column c (`col` maps 4-aligned groups of output columns to 4 consecutive matrix columns), `bcol(c)` its bias;
// with LN the input rows are normalised on the fly ((x - mu) * rs * g + b: the reference's rounding order).
// Thread (row group of 9, 4 consecutive columns): x reads are warp broadcasts, the warp's weight read at one k is one
// contiguous 512-byte row segment (L1/L2-resident).
template <bool LN, int MODE, typename Col, typename BCol>
__device__ void ca_gemm(const float* in, const float* mu, const float* rs, const float* __restrict__ g,
                        const float* __restrict__ b, const float* __restrict__ wt, int ldw, Col col, BCol bcol, int n,
                        float* out) {
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c0 = cg * 4, r0 = rg * kCaRowsPerThread;
  if (c0 < n) {
    float acc[kCaRowsPerThread][4];
#pragma unroll
    for (int r = 0; r < kCaRowsPerThread; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    const float* wp = wt + col(c0);
    float m[kCaRowsPerThread], s[kCaRowsPerThread];
    if (LN) {
#pragma unroll
      for (int r = 0; r < kCaRowsPerThread; ++r) m[r] = mu[r0 + r], s[r] = rs[r0 + r];
    }
#pragma unroll 2
    for (int k = 0; k < kCaD; k += 4) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(wp + (k + 0) * ldw));  // input k:     columns c0 .. c0+3
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(wp + (k + 1) * ldw));
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(wp + (k + 2) * ldw));
      const float4 a3 = __ldg(reinterpret_cast<const float4*>(wp + (k + 3) * ldw));
      float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (LN) {
        g4 = __ldg(reinterpret_cast<const float4*>(g + k));
        b4 = __ldg(reinterpret_cast<const float4*>(b + k));
      }
#pragma unroll
      for (int r = 0; r < kCaRowsPerThread; ++r) {
        float4 x = *reinterpret_cast<const float4*>(in + (r0 + r) * kCaLd + k);
        if (LN) {
          x.x = (x.x - m[r]) * s[r] * g4.x + b4.x;
          x.y = (x.y - m[r]) * s[r] * g4.y + b4.y;
          x.z = (x.z - m[r]) * s[r] * g4.z + b4.z;
          x.w = (x.w - m[r]) * s[r] * g4.w + b4.w;
        }
        acc[r][0] = fmaf(x.x, a0.x, fmaf(x.y, a1.x, fmaf(x.z, a2.x, fmaf(x.w, a3.x, acc[r][0]))));
        acc[r][1] = fmaf(x.x, a0.y, fmaf(x.y, a1.y, fmaf(x.z, a2.y, fmaf(x.w, a3.y, acc[r][1]))));
        acc[r][2] = fmaf(x.x, a0.z, fmaf(x.y, a1.z, fmaf(x.z, a2.z, fmaf(x.w, a3.z, acc[r][2]))));
        acc[r][3] = fmaf(x.x, a0.w, fmaf(x.y, a1.w, fmaf(x.z, a2.w, fmaf(x.w, a3.w, acc[r][3]))));
      }
    }
    const float bias[4] = {bcol(c0), bcol(c0 + 1), bcol(c0 + 2), bcol(c0 + 3)};
#pragma unroll
    for (int r = 0; r < kCaRowsPerThread; ++r) {
      float4* o = reinterpret_cast<float4*>(out + (r0 + r) * kCaLd + c0);
      float4 v = make_float4(acc[r][0] + bias[0], acc[r][1] + bias[1], acc[r][2] + bias[2], acc[r][3] + bias[3]);
      if (MODE == CA_STORE_GELU) v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
      if (MODE == CA_ACC) {
        const float4 p = *o;
        v = make_float4(p.x + v.x, p.y + v.y, p.z + v.z, p.w + v.w);
      }
      *o = v;
    }
  }
  __syncthreads();
}

// y[n_out] = W[n_out, n_in] x + b: one warp per output row (coalesced weight reads), x and y in shared memory
__device__ void ca_linear(const float* x, int n_in, const float* __restrict__ wt, const float* __restrict__ bias, int n_out,
                          float* y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < n_out; o += kCaWarps) {
    float acc = 0.f;
    for (int k = lane; k < n_in; k += 32) acc = fmaf(x[k], __ldg(wt + static_cast<int64_t>(o) * n_in + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) y[o] = acc + __ldg(bias + o);
  }
  __syncthreads();
}

// Multi-head self-attention block of a pre-norm layer: x += out_proj(MHA(LN(x))).  Two heads per pass: q | k | v of the
// pair (96 columns) go to `qkv`, one thread per (query, head) runs an online softmax over the 144 keys and writes its 16
// output channels to `o`; after four passes `o` holds the concatenated heads.
__device__ void ca_self_attention(float* x, float* qkv, float* o, const CaShared& S, const CaAttnW& W,
                                  const float* __restrict__ ln_w, const float* __restrict__ ln_b) {
  ca_row_stats(x, S.mu, S.rs);
  for (int pass = 0; pass < 4; ++pass) {
    const float* in_w = W.in_w;
    const float* in_b = W.in_b;
    auto row = [=](int c) { return (c >> 5) * kCaD + pass * 32 + (c & 31); };  // q / k / v block, pair offset, channel
    ca_gemm<true, CA_STORE>(x, S.mu, S.rs, ln_w, ln_b, in_w, 3 * kCaD, row, [=](int c) { return __ldg(in_b + row(c)); }, 96,
                            qkv);
    if (threadIdx.x < 2 * kCaT) {
      const int h = threadIdx.x / kCaT, i = threadIdx.x - h * kCaT;
      float q[kCaDh], acc[kCaDh];
      const float* qp = qkv + i * kCaLd + h * kCaDh;
#pragma unroll
      for (int d = 0; d < kCaDh; ++d) q[d] = qp[d] * 0.25f, acc[d] = 0.f;  // q scaled by 1/sqrt(16)
      float mx = -INFINITY, l = 0.f;
      for (int j = 0; j < kCaT; ++j) {
        const float* kp = qkv + j * kCaLd + 32 + h * kCaDh;
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < kCaDh; ++d) sc = fmaf(q[d], kp[d], sc);
        const float mn = fmaxf(mx, sc);
        const float corr = expf(mx - mn), p = expf(sc - mn);
        l = fmaf(l, corr, p);
        const float* vp = kp + 32;
#pragma unroll
        for (int d = 0; d < kCaDh; ++d) acc[d] = fmaf(acc[d], corr, p * vp[d]);
        mx = mn;
      }
      const float inv = 1.f / l;
      float* op = o + i * kCaLd + pass * 32 + h * kCaDh;
#pragma unroll
      for (int d = 0; d < kCaDh; ++d) op[d] = acc[d] * inv;
    }
    __syncthreads();
  }
  const float* ow = W.out_w;
  const float* ob = W.out_b;
  ca_gemm<false, CA_ACC>(o, nullptr, nullptr, nullptr, nullptr, ow, kCaD, [](int c) { return c; },
                         [=](int c) { return __ldg(ob + c); }, kCaD, x);
}

// Feed-forward block of a pre-norm layer: x += linear2(GELU(linear1(LN(x)))), hidden units in four chunks of 128.
__device__ void ca_feed_forward(float* x, float* hid, float* sum, const CaShared& S, const float* __restrict__ l1_w,
                                const float* __restrict__ l1_b, const float* __restrict__ l2_w,
                                const float* __restrict__ l2_b, const float* __restrict__ ln_w,
                                const float* __restrict__ ln_b) {
  ca_row_stats(x, S.mu, S.rs);
  for (int ch = 0; ch < kCaFf / kCaD; ++ch) {
    ca_gemm<true, CA_STORE_GELU>(x, S.mu, S.rs, ln_w, ln_b, l1_w, kCaFf, [=](int c) { return ch * kCaD + c; },
                                 [=](int c) { return __ldg(l1_b + ch * kCaD + c); }, kCaD, hid);
    const float* w2 = l2_w + ch * kCaD * kCaD;   // rows ch*128 .. of linear2^T [512, 128]
    auto ident = [](int c) { return c; };
    if (ch == 0)
      ca_gemm<false, CA_STORE>(hid, nullptr, nullptr, nullptr, nullptr, w2, kCaD, ident, [=](int c) { return __ldg(l2_b + c); },
                               kCaD, sum);
    else
      ca_gemm<false, CA_ACC>(hid, nullptr, nullptr, nullptr, nullptr, w2, kCaD, ident, [](int) { return 0.f; }, kCaD, sum);
  }
  for (int e = threadIdx.x; e < kCaT * (kCaD / 4); e += kCaThreads) {
    const int r = e >> 5, c = (e & 31) * 4;
    float4* xp = reinterpret_cast<float4*>(x + r * kCaLd + c);
    const float4 a = *xp, b = *reinterpret_cast<const float4*>(sum + r * kCaLd + c);
    *xp = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  __syncthreads();
}

// nn.GroupNorm(8, C, eps 1e-5) + GELU in place on a channel-major [C, hw] map: each group is `gsize` contiguous floats.
__device__ void ca_groupnorm_gelu(float* x, int channels, int hw, const float* __restrict__ w, const float* __restrict__ b,
                                  float* gstat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gsize = channels / 8 * hw;
  if (warp < 8) {
    const float* p = x + warp * gsize;
    float s = 0.f;
    for (int i = lane; i < gsize; i += 32) s += p[i];
    const float m = warp_sum(s) / static_cast<float>(gsize);
    float q = 0.f;
    for (int i = lane; i < gsize; i += 32) {
      const float d = p[i] - m;
      q = fmaf(d, d, q);
    }
    q = warp_sum(q) / static_cast<float>(gsize);
    if (lane == 0) gstat[2 * warp] = m, gstat[2 * warp + 1] = rsqrtf(q + 1e-5f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < channels * hw; i += kCaThreads) {
    const int c = i / hw, g = i / gsize;
    x[i] = gelu_erf((x[i] - gstat[2 * g]) * gstat[2 * g + 1] * __ldg(w + c) + __ldg(b + c));
  }
  __syncthreads();
}

// Conv2d(ci, co, 3, stride 2, padding 1): in [ci, h, w] -> out [co, h/2, w/2], both channel-major in shared memory
__device__ void ca_conv3x3s2(const float* in, int ci, int h, int w, const float* __restrict__ wt,
                             const float* __restrict__ bias, int co, float* out) {
  const int oh = h >> 1, ow = w >> 1, ohw = oh * ow, hw = h * w;
  for (int o = threadIdx.x; o < co * ohw; o += kCaThreads) {
    const int c = o / ohw, p = o - c * ohw, y = p / ow, x = p - y * ow;
    float acc = __ldg(bias + c);
    for (int k = 0; k < ci; ++k) {
      const float* wk = wt + (c * ci + k) * 9;
      const float* ik = in + k * hw;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yy = 2 * y - 1 + r;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int xx = 2 * x - 1 + s;
          if (xx < 0 || xx >= w) continue;
          acc = fmaf(ik[yy * w + xx], __ldg(wk + r * 3 + s), acc);
        }
      }
    }
    out[o] = acc;
  }
  __syncthreads();
}

// ConvTranspose2d(ci, co, 4, stride 2, padding 1): out [co, 2h, 2w] channel-major; weight [ci, co, 4, 4]. The input
// element (k, pixel p) lives at in[p * pix_stride + k * ch_stride] (tokens: 132 / 1, channel-major maps: 1 / h*w).
// Output row Y receives input rows y with Y = 2y - 1 + r: r has the parity of Y + 1.
__device__ void ca_convT4x4s2(const float* in, int pix_stride, int ch_stride, int ci, int h, int w,
                              const float* __restrict__ wt, const float* __restrict__ bias, int co, float* out) {
  const int oh = 2 * h, ow = 2 * w, ohw = oh * ow;
  for (int o = threadIdx.x; o < co * ohw; o += kCaThreads) {
    const int c = o / ohw, p = o - c * ohw, Y = p / ow, X = p - Y * ow;
    float acc = __ldg(bias + c);
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      const int r = ((Y + 1) & 1) + 2 * ri;
      const int y2 = Y + 1 - r;  // = 2y
      if (y2 < 0 || y2 >= 2 * h) continue;
#pragma unroll
      for (int si = 0; si < 2; ++si) {
        const int s = ((X + 1) & 1) + 2 * si;
        const int x2 = X + 1 - s;
        if (x2 < 0 || x2 >= 2 * w) continue;
        const float* ip = in + ((y2 >> 1) * w + (x2 >> 1)) * pix_stride;
        const float* wp = wt + c * 16 + r * 4 + s;
        for (int k = 0; k < ci; ++k) acc = fmaf(ip[k * ch_stride], __ldg(wp + k * co * 16), acc);
      }
    }
    out[o] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCaThreads, 1) convattn_kernel(const float* __restrict__ x, int cin, int layers,
                                                                int latent_dim, ConvAttnWeights W, float* __restrict__ z,
                                                                float* __restrict__ recon, double* __restrict__ huber_sums,
                                                                int do_encode, int do_decode) {
  extern __shared__ __align__(16) float s_ca[];
  CaShared S;
  S.buf[0] = s_ca;
  S.buf[1] = s_ca + kCaBuf;
  S.buf[2] = s_ca + 2 * kCaBuf;
  S.mu = s_ca + 3 * kCaBuf;
  S.rs = S.mu + kCaT;
  S.vec[0] = S.rs + kCaT;
  S.vec[1] = S.vec[0] + kCaD;
  S.vec[2] = S.vec[1] + kCaD;
  S.gstat = S.vec[2] + kCaD;
  float* s_red = S.gstat + 16;  // [16]
  float* X = S.buf[0];
  float* A = S.buf[1];
  float* B = S.buf[2];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* xin = x ? x + static_cast<int64_t>(n) * cin * 2304 : nullptr;
  float* zrow = z + static_cast<int64_t>(n) * latent_dim;

  if (do_encode) {
    // ---- encoder_cnn (train.py:69-76): frame -> B, conv -> arena [64, 24, 24] over X|A, conv -> B [128, 12, 12]
    for (int i = threadIdx.x; i < cin * 2304; i += kCaThreads) B[i] = __ldg(xin + i);
    __syncthreads();
    ca_conv3x3s2(B, cin, 48, 48, W.c0_w, W.c0_b, 64, X);
    ca_groupnorm_gelu(X, 64, 576, W.g0_w, W.g0_b, S.gstat);
    ca_conv3x3s2(X, 64, 24, 24, W.c1_w, W.c1_b, kCaD, B);
    ca_groupnorm_gelu(B, kCaD, kCaT, W.g1_w, W.g1_b, S.gstat);
    // flatten(2).transpose(1, 2) + encoder_pos_embedding (train.py:134-135)
    for (int e = threadIdx.x; e < kCaT * kCaD; e += kCaThreads) {
      const int p = e >> 7, c = e & 127;
      X[p * kCaLd + c] = B[c * kCaT + p] + __ldg(W.enc_pos + e);
    }
    __syncthreads();
    for (int l = 0; l < layers; ++l) {
      const CaEncLayerW& L = W.enc[l];
      ca_self_attention(X, A, B, S, L.sa, L.n1_w, L.n1_b);
      ca_feed_forward(X, A, B, S, L.l1_w, L.l1_b, L.l2_w, L.l2_b, L.n2_w, L.n2_b);
    }
    // ---- attention_pool: one query over the context (train.py:137-138); k -> A, v -> B, scores -> X (context is dead)
    {
      const float* in_w = W.pool.in_w;
      const float* in_b = W.pool.in_b;
      ca_linear(W.pool_q, kCaD, in_w, in_b, kCaD, S.vec[0]);  // q = Wq pooling_query + bq (global input is fine here)
      ca_gemm<false, CA_STORE>(X, nullptr, nullptr, nullptr, nullptr, W.pool_in_wt, 3 * kCaD, [](int c) { return kCaD + c; },
                               [=](int c) { return __ldg(in_b + kCaD + c); }, kCaD, A);
      ca_gemm<false, CA_STORE>(X, nullptr, nullptr, nullptr, nullptr, W.pool_in_wt, 3 * kCaD,
                               [](int c) { return 2 * kCaD + c; }, [=](int c) { return __ldg(in_b + 2 * kCaD + c); }, kCaD, B);
      float* sc = X;  // [8, 144]
      for (int e = threadIdx.x; e < 8 * kCaT; e += kCaThreads) {
        const int h = e / kCaT, j = e - h * kCaT;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < kCaDh; ++k) d = fmaf(S.vec[0][h * kCaDh + k] * 0.25f, A[j * kCaLd + h * kCaDh + k], d);
        sc[e] = d;
      }
      __syncthreads();
      if (warp < 8) {
        float* row = sc + warp * kCaT;
        float mx = -INFINITY;
        for (int j = lane; j < kCaT; j += 32) mx = fmaxf(mx, row[j]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < kCaT; j += 32) {
          const float p = expf(row[j] - mx);
          row[j] = p;
          sum += p;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int j = lane; j < kCaT; j += 32) row[j] *= inv;
      }
      __syncthreads();
      if (threadIdx.x < kCaD) {
        const int c = threadIdx.x, h = c / kCaDh;
        float acc = 0.f;
        for (int j = 0; j < kCaT; ++j) acc = fmaf(sc[h * kCaT + j], B[j * kCaLd + c], acc);
        S.vec[1][c] = acc;
      }
      __syncthreads();
      ca_linear(S.vec[1], kCaD, W.pool.out_w, W.pool.out_b, kCaD, S.vec[0]);  // pooled
      // encoder_head: LayerNorm(128) + Linear(128 -> latent) (train.py:93-96, 139)
      if (warp == 0) {
        const float4 v = *reinterpret_cast<const float4*>(S.vec[0] + lane * 4);
        const float m = warp_sum(v.x + v.y + v.z + v.w) * (1.f / kCaD);
        const float a = v.x - m, b = v.y - m, c = v.z - m, d = v.w - m;
        const float r = rsqrtf(warp_sum(a * a + b * b + c * c + d * d) * (1.f / kCaD) + 1e-5f);
        const float4 g = __ldg(reinterpret_cast<const float4*>(W.eh_ln_w) + lane);
        const float4 be = __ldg(reinterpret_cast<const float4*>(W.eh_ln_b) + lane);
        *reinterpret_cast<float4*>(S.vec[1] + lane * 4) =
            make_float4(a * r * g.x + be.x, b * r * g.y + be.y, c * r * g.z + be.z, d * r * g.w + be.w);
      }
      __syncthreads();
      ca_linear(S.vec[1], kCaD, W.eh_w, W.eh_b, latent_dim, A);  // z in A[0 .. latent)
      for (int i = threadIdx.x; i < latent_dim; i += kCaThreads) zrow[i] = A[i];
      __syncthreads();
    }
  } else {
    for (int i = threadIdx.x; i < latent_dim; i += kCaThreads) A[i] = __ldg(zrow + i);
    __syncthreads();
  }
  if (!do_decode) return;

  // ---- decode (train.py:142-156): memory token, queries, L decoder layers, decoder_cnn
  ca_linear(A, latent_dim, W.dh_w, W.dh_b, kCaD, S.vec[2]);  // context [128]
  for (int e = threadIdx.x; e < kCaT * kCaD; e += kCaThreads)
    X[(e >> 7) * kCaLd + (e & 127)] = __ldg(W.dec_q + e) + __ldg(W.dec_pos + e);
  __syncthreads();
  for (int l = 0; l < layers; ++l) {
    const CaDecLayerW& L = W.dec[l];
    ca_self_attention(X, A, B, S, L.sa, L.n1_w, L.n1_b);
    // cross-attention over the single memory token: x += out_proj(v_proj(context))
    ca_linear(S.vec[2], kCaD, L.ca.in_w + 2 * kCaD * kCaD, L.ca.in_b + 2 * kCaD, kCaD, S.vec[0]);
    ca_linear(S.vec[0], kCaD, L.ca.out_w, L.ca.out_b, kCaD, S.vec[1]);
    for (int e = threadIdx.x; e < kCaT * kCaD; e += kCaThreads) X[(e >> 7) * kCaLd + (e & 127)] += S.vec[1][e & 127];
    __syncthreads();
    ca_feed_forward(X, A, B, S, L.l1_w, L.l1_b, L.l2_w, L.l2_b, L.n3_w, L.n3_b);
  }
  // patches.transpose(1, 2).reshape(b, 128, 12, 12): channel k of pixel p is token p, feature k (train.py:154)
  float* mid = A;  // [64, 24, 24] over A|B
  ca_convT4x4s2(X, kCaLd, 1, kCaD, 12, 12, W.t0_w, W.t0_b, 64, mid);
  ca_groupnorm_gelu(mid, 64, 576, W.g2_w, W.g2_b, S.gstat);
  ca_convT4x4s2(mid, 1, 576, 64, 24, 24, W.t1_w, W.t1_b, cin, X);
  float hub = 0.f;
  for (int i = threadIdx.x; i < cin * 2304; i += kCaThreads) {
    const float v = X[i];
    recon[static_cast<int64_t>(n) * cin * 2304 + i] = v;
    if (xin != nullptr) {
      const float d = fabsf(v - __ldg(xin + i));
      hub += d < 1.f ? 0.5f * d * d : d - 0.5f;  // nn.HuberLoss(delta = 1)
    }
  }
  if (huber_sums != nullptr && xin != nullptr) {
    hub = warp_sum(hub);
    if (lane == 0) s_red[warp] = hub;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < kCaWarps; ++i) tot += s_red[i];
      atomicAdd(&huber_sums[0], static_cast<double>(tot));
      atomicAdd(&huber_sums[1], static_cast<double>(cin * 2304));
    }
  }
}

}  // namespace wfk

// weights: 29 + 30 L pointers (L = layers) in the order of the reference module (see
// weatherforecastingtoolkit_b200/predictors.py::ConvAttnModel._weight_pointers).
extern "C" int wfk_convattn_forward(const float* x, int n, int cin, int layers, int latent_dim, const float* const* weights,
                                    int num_weights, float* z, float* recon, double* huber_sums, int mode, void* stream) {
  WFK_ENTER(stream, (x != nullptr ? static_cast<const void*>(x) : static_cast<const void*>(z)));
  WFK_REQUIRE(weights && z, "null pointer");
  WFK_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (encode + decode), 1 (encode only) or 2 (decode only)");
  WFK_REQUIRE(mode == 2 || x != nullptr, "x is required unless decoding from z");
  WFK_REQUIRE(mode == 1 || recon != nullptr, "recon is required unless encoding only");
  WFK_REQUIRE(n > 0 && cin >= 1 && cin <= 8, "bad shape n=%d cin=%d", n, cin);
  WFK_REQUIRE(layers >= 1 && layers <= wfk::kCaMaxLayers, "1..%d transformer layers supported", wfk::kCaMaxLayers);
  WFK_REQUIRE(latent_dim >= 4 && latent_dim <= wfk::kCaMaxLatent, "latent_dim must be in 4..%d", wfk::kCaMaxLatent);
  const int expect = 29 + 30 * layers;
  WFK_REQUIRE(num_weights == expect, "expected %d weight pointers for %d layers, got %d", expect, layers, num_weights);
  for (int i = 0; i < num_weights; ++i) {
    WFK_REQUIRE(weights[i] != nullptr, "weights[%d] is NULL", i);
    WFK_REQUIRE((reinterpret_cast<uintptr_t>(weights[i]) & 15) == 0, "weights[%d] is not 16-byte aligned", i);
  }
  wfk::ConvAttnWeights W = {};
  int k = 0;
  auto next = [&]() { return weights[k++]; };
  auto attn = [&](wfk::CaAttnW& a) { a.in_w = next(), a.in_b = next(), a.out_w = next(), a.out_b = next(); };
  W.c0_w = next(), W.c0_b = next(), W.g0_w = next(), W.g0_b = next();
  W.c1_w = next(), W.c1_b = next(), W.g1_w = next(), W.g1_b = next();
  W.enc_pos = next();
  for (int l = 0; l < layers; ++l) {
    wfk::CaEncLayerW& L = W.enc[l];
    attn(L.sa);
    L.l1_w = next(), L.l1_b = next(), L.l2_w = next(), L.l2_b = next();
    L.n1_w = next(), L.n1_b = next(), L.n2_w = next(), L.n2_b = next();
  }
  W.pool_q = next();
  attn(W.pool);
  W.pool_in_wt = next();
  W.eh_ln_w = next(), W.eh_ln_b = next(), W.eh_w = next(), W.eh_b = next();
  W.dh_w = next(), W.dh_b = next();
  W.dec_q = next(), W.dec_pos = next();
  for (int l = 0; l < layers; ++l) {
    wfk::CaDecLayerW& L = W.dec[l];
    attn(L.sa);
    attn(L.ca);
    L.l1_w = next(), L.l1_b = next(), L.l2_w = next(), L.l2_b = next();
    L.n1_w = next(), L.n1_b = next(), L.n2_w = next(), L.n2_b = next(), L.n3_w = next(), L.n3_b = next();
  }
  W.t0_w = next(), W.t0_b = next(), W.g2_w = next(), W.g2_b = next(), W.t1_w = next(), W.t1_b = next();
  const size_t smem = (static_cast<size_t>(3) * wfk::kCaBuf + 2 * wfk::kCaT + 3 * wfk::kCaD + 32) * sizeof(float);
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::convattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    attr_lock.finished();
  }
  wfk::convattn_kernel<<<n, wfk::kCaThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, cin, layers, latent_dim, W, z, recon, huber_sums, mode != 2, mode != 1);
  return wfk::launched("convattn_kernel");
}
