// Implicit-GEMM convolution / GEMM for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
// Data layout: activations NHWC fp16 (channels innermost) so that a block of 128 output pixels x 64
// input channels is a TMA box (64, BW, 1, BH, 1) of the 5-D view (k, x, q, y, frame); TMA writes it
// as 128 rows of 128 B with the 128-byte swizzle, which is exactly the K-major operand layout
// tcgen05.mma reads. A 3x3 convolution is 9 "taps": the same box shifted by (dx, dy); rows that fall
// outside the image are zero-filled by TMA, which implements the convolution's zero padding.
// Weights are [slab][cout][cin] fp16 and arrive as the box (64, rows, 1) of the 3-D view (k, n, slab).
//
// Warp roles (320 threads, one persistent CTA per SM):
//   warp 0      : TMA producer (one elected lane)       -- fills the STAGES-deep smem ring
//   warp 1      : TMEM allocator + MMA issuer (one lane) -- 4 x tcgen05.mma (K=16) per 64-wide K block
//   warps 2..9  : epilogue -- tcgen05.ld the fp32 accumulator, + bias (+ residual), GroupNorm
//                 sum / sum-of-squares of the result, fp16 / fp32 stores (two warps per TMEM lane quarter)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Tile shapes: BN=256: one 128-pixel block per CTA; BN=128: MB=2 pixel blocks per CTA share the weight
// tile. PAIR=true runs the kernel as 2-CTA clusters with tcgen05.mma.cta_group::2 (M=256 across the
// pair): each CTA stages its own pixel blocks and only HALF of the weight tile, which cuts the shared
// memory traffic per MAC by a third -- the limiter of the single-CTA kernel.
//
// Replaces (reference, all cuDNN / cuBLAS library calls): nn.Conv2d 3x3 / 1x1
// (pipeline/models/autoencoderkl/resnet.py:405,421,452), Downsample2D (resnet.py:181-190),
// Upsample2D (resnet.py:108-143), the attention linears and bmm's (attention.py:146-176).
#include <cstdlib>
#include <new>

#include "act16.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace wfk {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // fp16 elements = one 128 B swizzle row
constexpr int kEpiWarps = 8;  // two per TMEM lane quarter: latency hiding by TLP
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreadsPlain = 64 + kEpiThreads;   // warps: TMA, MMA, 8 epilogue
// HALO: GroupNorm+SiLU applied in place to the staged halo by 4 (MB = 1) or 8 (MB = 2) transform warps. The MB = 2
// kernel (Cout = 128 layers at 384^2) stages 324 halo rows per 64-channel K block against the same 4608 tensor-core
// cycles as the MB = 1 kernel's 180 rows: with 4 warps the transform, not the tensor pipe, bounded those layers.
// 8 transform + 8 epilogue + 4 (TMA-B, TMA-A, MMA, idle) warps = 640 threads only fit the register file because the
// roles trade registers with setmaxnreg (warpgroup-aligned roles: transform 64, epilogue 144, producers / MMA 56:
// 8*64 + 8*144 + 4*56 = 1888 <= 20*96 -- setmaxnreg.inc draws only on registers the CTA's own warps released).
__host__ __device__ constexpr int xform_warps(bool halo, int mb) { return halo ? (mb == 2 ? 8 : 4) : 0; }
__host__ __device__ constexpr int conv_threads(bool halo, int mb) {
  return halo ? (mb == 2 ? 640 : 96 + 32 * 4 + kEpiThreads) : kConvThreadsPlain;
}
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
constexpr int kABytes = kBlockM * kBlockK * 2;
// HALO mode (3x3 stride-1 taps): pixel blocks are 8 wide x 16 tall; ONE (8*MB+2) x 18 pixel halo box per
// 64-channel K block serves all taps through shifted A descriptors (tcgen05's 128B swizzle is a pure
// function of the shared-memory address -- probed in csrc/debug_mma.cu), so activations are staged once
// instead of once per tap and the weight tiles stream through their own ring.
constexpr int kHaloRows = 18;
constexpr int kHaloBStagesMax = 8;

// Timing experiments (operand traffic removed, stores skipped, fewer CTA pairs ...) produce WRONG results by design:
// they exist only in builds made with -DWFK_EXPERIMENTS and can never be switched on in the product library.
#ifdef WFK_EXPERIMENTS
#define WFK_XDBG(p) ((p).xform_debug)
#define WFK_HINT_NS(p) ((p).wait_hint_ns)
#else
#define WFK_XDBG(p) 0
#define WFK_HINT_NS(p) 0
#endif

struct ConvKernelParams {
  CUtensorMap a_map[2];
  CUtensorMap b_map[2];
  int n_frames, tile_h, tile_w, tiles_x, tiles_y, tiles_n, n_total;
  int bw_log2;
  int stack_x;  // pixel blocks of one tile are laid along x (GEMM-like, one row) instead of y
  int seg_kblocks[2];  // HALO: K blocks of source 0 (all taps) and of source 1 (fused 1x1 shortcut, centre tap)
  int seg1_slab;       // HALO: weight slab of the shortcut
  const float2* gn_table;  // HALO: [frame][cin] (scale, shift) of a fused GroupNorm+SiLU on source 0, or null
  int gn_cin;
  // HALO, alternative to gn_table: the (scale, shift) pairs are derived inside the kernel from the producer's raw
  // statistics, which removes one tiny kernel launch (and its pipeline bubble) in front of every fused convolution
  const double* gn_stats;  // [frame][gn_groups][2] (sum, sum of squares)
  const float* gn_gamma;
  const float* gn_beta;
  double gn_inv_count;     // 1 / (channels per group * pixels)
  float gn_eps;
  int gn_cpg_log2, gn_groups;
  int wait_hint_ns; // >0: epilogue / producer waits park with this try_wait suspend hint
  int xform_debug;  // experiment switch: 1 = load/store without math, 2 = skip the transform entirely
  int num_phases, taps_per_phase;
  int a_frame_mul, b_frame_mul;
  const float* bias;
  const __half* residual;
  __half* out_h;
  float* out_f;
  double* stats;
  int out_rows, out_cols, out_sy, out_sx, ldc;
  int cpg_log2, groups_total;
  // EPI kernels only: v = act(acc + bias (+ residual)) -> out_h / out_f;  out2_h = act2(scale2[c] * v + shift2[c])
  int act, act2;
  float act_slope;
  __half* out2_h;
  const float* scale2;
  const float* shift2;
  // EPI kernels without HALO: row-softmax passes (WFK_ACT_ROW_MAX / ROW_EXP / ROW_NORM), [rows][row_ld] fp32 partials
  const float* row_in;
  float* row_out;
  int row_ld;
  float row_scale;
  uint32_t idesc;
  // Cout = 128 HALO kernel (MB = 2): 16-bit output through TMA stores. (c, x, y, frame) view of the NHWC output, box
  // 32 channels x 8 x 4 pixels = one epilogue warp's chunk, 64-byte swizzle = the warp's staging-tile layout.
  int use_tma_store;
  CUtensorMap out_map;
  wfk_tap taps[WFK_MAX_TAPS];
};

__device__ __forceinline__ void mbar_wait_h(uint64_t* bar, uint32_t parity, int hint_ns) {
  if (hint_ns > 0) mbar_wait_parked(bar, parity, static_cast<uint32_t>(hint_ns));
  else mbar_wait(bar, parity);
}

struct TileCoord {
  int phase, frame, ty, tx, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int tile) {
  TileCoord t;
  const int per_frame = p.tiles_y * p.tiles_x;
  const int per_phase = p.n_frames * per_frame * p.tiles_n;
  t.phase = tile / per_phase;
  int r = tile - t.phase * per_phase;
  const int mt = r / p.tiles_n;
  t.nt = r - mt * p.tiles_n;
  t.frame = mt / per_frame;
  r = mt - t.frame * per_frame;
  t.ty = r / p.tiles_x;
  t.tx = r - t.ty * p.tiles_x;
  return t;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Epilogue activations of the EPI kernels (wfk_act in the C ABI).
__device__ __forceinline__ float act_apply(int kind, float x, float slope) {
  switch (kind) {
    case WFK_ACT_LEAKY_RELU: return x > 0.f ? x : slope * x;
    case WFK_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));  // nn.GELU() default (erf form)
    case WFK_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case WFK_ACT_SILU: return x / (1.f + __expf(-x));
    default: return x;
  }
}

// Sum / sum-of-squares of 32 consecutive channels (one accumulator row chunk per lane) reduced over
// the 32 lanes (pixels) of the warp for each of NG = 32/cpg groups, with a transposing butterfly:
// after log2(NG) exchange steps every lane owns one group, then a plain xor-reduction finishes.
template <int NG>
__device__ __forceinline__ void chunk_group_stats(const float (&v)[32], bool valid, int lane, float* s_dst) {
  constexpr int CPG = 32 / NG;
  float s[NG], q[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float x = v[g * CPG + j];
      a += x;
      b = fmaf(x, x, b);
    }
    s[g] = valid ? a : 0.f;
    q[g] = valid ? b : 0.f;
  }
  int grp = 0;
#pragma unroll
  for (int n = NG, off = 16; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
    grp += upper ? (n >> 1) : 0;
#pragma unroll
    for (int i = 0; i < (n >> 1); ++i) {
      const float ks = upper ? s[i + (n >> 1)] : s[i];
      const float ss = upper ? s[i] : s[i + (n >> 1)];
      const float kq = upper ? q[i + (n >> 1)] : q[i];
      const float sq = upper ? q[i] : q[i + (n >> 1)];
      s[i] = ks + __shfl_xor_sync(0xffffffffu, ss, off);
      q[i] = kq + __shfl_xor_sync(0xffffffffu, sq, off);
    }
  }
  constexpr int REM = 32 / NG;  // lanes that still hold partials of the same group
#pragma unroll
  for (int off = REM >> 1; off >= 1; off >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], off);
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], off);
  }
  if ((lane & (REM - 1)) == 0) {  // s_dst is this warp's private slot: plain RMW, fixed order
    s_dst[2 * grp + 0] += s[0];
    s_dst[2 * grp + 1] += q[0];
  }
}

// Transposing warp reduction: every lane holds 32 partial values; afterwards lane L holds the warp-wide total of the
// value whose index is returned (31 shuffles for 32 x 32 partials instead of 160 for 32 separate xor reductions).
__device__ __forceinline__ int warp_transpose_reduce32(float (&v)[32], int lane) {
  int idx = 0;
#pragma unroll
  for (int n = 32, off = 16; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
    idx += upper ? (n >> 1) : 0;
#pragma unroll
    for (int i = 0; i < (n >> 1); ++i) {
      const float keep = upper ? v[i + (n >> 1)] : v[i];
      const float send = upper ? v[i] : v[i + (n >> 1)];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return idx;
}

template <int BN, int MB, int STAGES, bool PAIR, bool HALO, bool EPI, bool BF16>
__global__ void __launch_bounds__(conv_threads(HALO, MB), 1)
    conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  constexpr int kXformWarps = xform_warps(HALO, MB);
  constexpr bool kRegSplit = HALO && MB == 2;      // 20 warps: per-role register budgets via setmaxnreg
  // Row-contiguous (shared-memory staged) epilogue accesses: pays for the 128-wide tiles, whose K = 9*128 layers are
  // epilogue / LSU-bound; the 256-wide tiles keep direct row-per-lane accesses (their epilogue has 2-4x the MMA time
  // to hide in, and the extra staging traffic and the shallower operand ring cost more than the stores)
  constexpr bool kCoalesce = true;
  // The 256-wide HALO kernel has ~10 KB of shared memory to spare: its staging tile holds 16 rows and a chunk goes
  // through it in two passes (rows 0-15, rows 16-31); every other kernel stages all 32 rows at once.
  constexpr int kStPasses = (HALO && BN == 256) ? 2 : 1;
  constexpr int kStRows = 32 / kStPasses;
  constexpr int kStAcc = 4 / kStPasses;            // coalesced accesses (8 rows x 64 B each) per pass
  // fp32 outputs (attention scores) of the plain-tap kernels: the same staging with 128-byte rows
  constexpr bool kCoalesceF32 = !HALO;
  // The Cout = 128 layers at 384^2 are epilogue-bound (K = 9*128 gives the epilogue 0.28 tensor cycles per output) and
  // `ncu` shows the epilogue warps waiting on shared-memory loads (the staged tile read back for the row-contiguous
  // stores: short-scoreboard stalls, 14 % of their time, under the MMA's operand traffic). This kernel therefore hands
  // the staged 32-pixel x 32-channel chunk to the TMA engine instead: 4 x STS, one fence, ONE cp.async.bulk.tensor by an
  // elected lane -- no LDS, no STG, no per-lane output addressing, image-edge clipping by the tensor map.
  // The 256-wide HALO kernel does the same with its 16-row staging tile (two 1 KB stores per chunk): its epilogue is not
  // critical, but the LDS / STG / address work it no longer does is shared-memory bandwidth and power returned to the MMAs.
  constexpr bool kTmaStore = HALO && !EPI;
  constexpr bool kRowModes = EPI && !HALO;         // row-softmax epilogues (attention GEMMs)
  constexpr uint32_t kStageWarpBytes = HALO ? 2048u / kStPasses : 4096u;
  constexpr int kBRows = PAIR ? BN / 2 : BN;       // weight rows this CTA stages
  constexpr int kBBytes = kBRows * kBlockK * 2;
  constexpr int kCtas = PAIR ? 2 : 1;
  constexpr int kBlocksPerTile = MB * kCtas;       // 128-pixel blocks per (pair-)tile
  // HALO: A ring of STAGES halo boxes + B ring of kHaloBStages weight tiles; else one ring of (A blocks + B)
  constexpr int kHaloPitch = 8 * MB + 2;           // halo box width in pixels (= rows of 128 B per image row)
  constexpr int kHaloBytes = kHaloPitch * kHaloRows * 128;
  constexpr int kHaloStage = (kHaloBytes + 1023) & ~1023;
  // halo transform: 128 threads = 16 row slots x 8 sixteen-byte chunks; kXfPasses row passes per box, the loads of
  // kXfGroup passes in flight per thread (MB=2: 324 rows -> 21 = 3 x 7 passes; MB=1: 180 rows -> 12 = 2 x 6)
  constexpr int kXfRowsPerPass = HALO ? 4 * kXformWarps : 16;  // 8 threads (16-byte chunks) per 128-byte row
  constexpr int kXfPasses = (kHaloPitch * kHaloRows + kXfRowsPerPass - 1) / kXfRowsPerPass;
  constexpr int kXfGroup = 6;                      // loads in flight per thread (MB=2: 11 passes = 6 + 5; MB=1: 12 = 6 + 6)
  constexpr int kStageBytes = HALO ? kHaloStage : MB * kABytes + kBBytes;
  constexpr int kBStages = HALO ? (MB == 1 ? (EPI ? 5 : 6) : 5) : 0;  // weight-tile ring depth (16 KB / 8 KB tiles; the EPI
                                                                      // variant gives one up for its scale2 / shift2 arrays)
  // Warp roles by warp id. The SM's warp arbiter favours HIGHER warp ids, so the latency-critical
  // single-thread roles sit at the top: (HALO: transform 0..3,) epilogue (8 warps), (HALO: weight-tile
  // TMA,) activation TMA, then the MMA issuer last. The epilogue outranks the transform: with K = 9*128
  // it is co-critical with the MMAs, whereas a halo transform has nine taps' worth of slack.
  // (Measured: for the 256-wide tiles the opposite order of these two is ~3 % faster, so it depends on MB.)
  // MB = 2 (kRegSplit): transform 0..7 | epilogue 8..15 | TMA-B 16, TMA-A 17, MMA 18, idle 19 (whole warpgroups per role)
  constexpr int kXformWarp0 = (MB == 2) ? 0 : kEpiWarps;
  constexpr int kFirstEpiWarp = (HALO && MB == 2) ? kXformWarps : 0;
  constexpr int kWarpTmaB = HALO ? kEpiWarps + kXformWarps : -1;
  constexpr int kWarpTmaA = HALO ? kEpiWarps + kXformWarps + 1 : kEpiWarps;
  constexpr int kWarpMma = kWarpTmaA + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + STAGES * kStageBytes;   // HALO only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kBStages * kBBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* ready_bar = empty_bar + STAGES;        // HALO only: halo transformed (or passed through)
  uint64_t* bfull_bar = ready_bar + (HALO ? STAGES : 0);  // HALO only (kBStages each)
  uint64_t* bempty_bar = bfull_bar + kBStages;
  uint64_t* tfull_bar = bempty_bar + kBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_stats = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));  // [kEpiWarps][BN/4 groups max][2], 16 B aligned (s_bias is read as float4)
  float* s_bias = s_stats + kEpiWarps * (BN / 2);            // [BN] bias of the current N tile
  float* s_sc2 = s_bias + BN;                                // EPI only: [BN] scale2, [BN] shift2 of the N tile
  float* s_sh2 = s_sc2 + BN;
  // per-epilogue-warp staging tile (32 rows x 64 B, 16-byte units XOR-swizzled): turns the row-per-lane register
  // layout of a TMEM chunk into row-contiguous global accesses
  // (1 KB-aligned: the TMA store's 64-byte swizzle is a function of the shared-memory address bits 7-8)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(s_bias + BN * (EPI ? 3 : 1)) + (kTmaStore ? 1023 : 0)) & ~uintptr_t(kTmaStore ? 1023 : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 32 * kWarpMma) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);       // PAIR: only the leader arrives (expect_tx of BOTH CTAs' bytes)
      mbar_init(&empty_bar[i], 1);
    }
    if (HALO)
      for (int i = 0; i < STAGES; ++i) mbar_init(&ready_bar[i], kCtas * kXformWarps);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bfull_bar[i], 1);
      mbar_init(&bempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kCtas * kEpiWarps);  // one arrive per epilogue warp (PAIR: of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == kWarpTmaA && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.b_map[0]);
  }
  if (warp == kWarpMma) {
    if (PAIR) tmem_alloc_2sm<2 * MB * BN>(tmem_slot);
    else tmem_alloc<2 * MB * BN>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
  // the tail of the previous kernel in the stream; nothing below touches global memory before that kernel has
  // completed and flushed. The trigger lets the NEXT kernel's CTAs take over SMs as this grid's CTAs retire.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int total_tiles = p.num_phases * p.n_frames * p.tiles_y * p.tiles_x * p.tiles_n;
  const int tile0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int bw = 1 << p.bw_log2;
  const int bh = kBlockM >> p.bw_log2;

  // K-block schedule of a HALO tile, shared by every role: the K blocks of source 1 (fused 1x1 shortcut: ONE tap per
  // staged block, 9x shorter than a halo stage) are spread evenly between those of source 0 instead of running back to
  // back at the end of the tile. Back to back they drained the operand ring faster than TMA + the halo transform could
  // refill it for the next tile (~1 800 idle tensor-pipe cycles per tile: the shortcut layers ran 15 % below their
  // siblings).
  auto run_steps = [&](auto&& step) {
    const int n0 = p.seg_kblocks[0], n1 = p.seg_kblocks[1];
    int j1 = 0;
    for (int i0 = 0; i0 < n0; ++i0) {
      step(0, i0);
      const int end = (n1 * (i0 + 1)) / n0;
      for (; j1 < end; ++j1) step(1, j1);
    }
  };
  auto role_tma_a = [&]() {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      if (HALO) {
        // one halo box per (tile, source, 64-channel block): pixels [x0-1, x0+8*MB] x [y0-1, y0+16]
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
          const TileCoord t = decode_tile(p, tile);
          const int x0 = (t.tx * kBlocksPerTile + static_cast<int>(cta_rank) * MB) * 8;
          const int y0 = t.ty * 16;
          // K blocks of source 0 (every tap) interleaved with those of source 1 (fused 1x1 shortcut, centre tap): run_steps
          auto step = [&](const int seg, const int kb) {
            const CUtensorMap* am = &p.a_map[seg];
            
              mbar_wait_h(&empty_bar[stage], phase ^ 1u, WFK_HINT_NS(p));
              uint8_t* sa = smem + stage * kStageBytes;
              // each CTA's box completes on its OWN barrier: its transform warps consume it first.
              // Source 1 (1x1 shortcut, centre tap only) needs no halo: a plain (8*MB) x 16 pixel box.
              if (seg == 0) {
                mbar_arrive_expect_tx(&full_bar[stage], kHaloBytes);
                tma_load_5d(sa, am, &full_bar[stage], kb * kBlockK, x0 - 1, 0, y0 - 1, t.frame * p.a_frame_mul);
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], MB * kABytes);
                tma_load_5d(sa, am, &full_bar[stage], kb * kBlockK, x0, 0, y0, t.frame * p.a_frame_mul);
              }
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            };
          run_steps(step);
        }
      } else
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord t = decode_tile(p, tile);
        // first pixel block of this CTA inside the tile
        const int blk0 = static_cast<int>(cta_rank) * MB;
        const int bx0 = p.stack_x ? t.tx * kBlocksPerTile + blk0 : t.tx;
        const int by0 = p.stack_x ? t.ty : t.ty * kBlocksPerTile + blk0;
        for (int ti = 0; ti < p.taps_per_phase; ++ti) {
          const wfk_tap tap = p.taps[t.phase * p.taps_per_phase + ti];
          const CUtensorMap* am = &p.a_map[tap.src];
          const CUtensorMap* bm = &p.b_map[tap.src];
          for (int kb = 0; kb < tap.kblocks; ++kb) {
            mbar_wait_h(&empty_bar[stage], phase ^ 1u, WFK_HINT_NS(p));
            uint8_t* sa = smem + stage * kStageBytes;
            if (PAIR) {
              // The peer's bytes complete_tx on the leader's barrier too; the peer itself never arrives (a
              // release.cluster arrive per K block costs a fence on the producer's critical path). A peer
              // running ahead only drives the tx-count negative until the leader's expect_tx lands.
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const int x0 = (p.stack_x ? bx0 + mb : bx0) * bw;
              const int y0 = (p.stack_x ? by0 : by0 + mb) * bh;
              if (PAIR)
                tma_load_5d_2sm(sa + mb * kABytes, am, &full_bar[stage], tap.c_off + kb * kBlockK, x0 + tap.dx, tap.q,
                                y0 + tap.dy, t.frame * p.a_frame_mul);
              else
                tma_load_5d(sa + mb * kABytes, am, &full_bar[stage], tap.c_off + kb * kBlockK, x0 + tap.dx, tap.q,
                            y0 + tap.dy, t.frame * p.a_frame_mul);
            }
            if (PAIR)
              tma_load_3d_2sm(sa + MB * kABytes, bm, &full_bar[stage], kb * kBlockK,
                              t.nt * BN + static_cast<int>(cta_rank) * kBRows, tap.b_slab + t.frame * p.b_frame_mul);
            else
              tma_load_3d(sa + MB * kABytes, bm, &full_bar[stage], kb * kBlockK, t.nt * BN,
                          tap.b_slab + t.frame * p.b_frame_mul);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  };
  auto role_mma = [&]() {
    if (leader) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only when PAIR)
      // The whole warp runs this warp-uniform loop (waits included); one elected lane issues the tcgen05
      // instructions. The loop must sustain one tcgen05.mma per 64 tensor-core cycles (N = 128), so descriptors
      // are built incrementally: the high word (SBO, version, swizzle) is constant, the low word is the stage's
      // start address >> 4 plus (bytes >> 4) per K slice / pixel block.
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t lo0 = static_cast<uint32_t>(desc0);
      const uint32_t idesc = p.idesc;
      if (HALO) {
        // A descriptors: start = halo base + (r*pitch + 8*mb + s) rows, 8-row groups `pitch` rows apart
        const uint32_t a_hi0 = (desc_hi & ~0x3FFFu) | static_cast<uint32_t>((kHaloPitch * 128) >> 4);
        const uint32_t a_hi1 = (desc_hi & ~0x3FFFu) | static_cast<uint32_t>((8 * MB * 128) >> 4);  // shortcut box pitch
        const uint32_t b_lo0 = lo0 + ((STAGES * kStageBytes) >> 4);
        int bstage = 0;
        uint32_t bphase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
          const TileCoord t = decode_tile(p, tile);
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MB * BN);
          uint32_t accumulate = 0;
          // K blocks of source 0 (every tap) interleaved with those of source 1 (fused 1x1 shortcut, centre tap): run_steps
          auto step = [&](const int seg, const int kb) {
            const int ntaps = seg == 0 ? p.taps_per_phase : 1;
            const uint32_t a_hi = seg == 0 ? a_hi0 : a_hi1;
            mbar_wait(&ready_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_lo = lo0 + static_cast<uint32_t>(stage) * (kStageBytes >> 4);
            for (int ti = 0; ti < ntaps; ++ti) {
              int r = 0, sx = 0;  // source 1: the box starts at the block origin
              if (seg == 0) {
                const wfk_tap tap = p.taps[t.phase * p.taps_per_phase + ti];
                r = tap.dy + 1;
                sx = tap.dx + 1;
              }
              mbar_wait(&bfull_bar[bstage], bphase);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(bstage) * (kBBytes >> 4);
              const uint32_t a_tap = a_lo + static_cast<uint32_t>((r * kHaloPitch + sx) * 8);  // rows * 128 B >> 4
              if (issuer) {
                // the 4 K slices of one pixel block are issued back to back (same accumulator)
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k);
                    const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | (a_tap + static_cast<uint32_t>(mb * 64 + 2 * k));
                    if (PAIR) umma_f16_2sm(d_tmem + mb * BN, adesc, bdesc, idesc, (k > 0) ? 1u : accumulate);
                    else umma_f16(d_tmem + mb * BN, adesc, bdesc, idesc, (k > 0) ? 1u : accumulate);
                  }
                }
                if (PAIR) umma_commit_2sm(&bempty_bar[bstage]);
                else umma_commit(&bempty_bar[bstage]);
              }
              __syncwarp();
              accumulate = 1;
              if (++bstage == kBStages) {
                bstage = 0;
                bphase ^= 1u;
              }
            }
            if (issuer) {
              if (PAIR) umma_commit_2sm(&empty_bar[stage]);
              else umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          };
          run_steps(step);
          if (issuer) {
            if (PAIR) umma_commit_2sm(&tfull_bar[acc]);
            else umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord t = decode_tile(p, tile);
        int kb_total = 0;
        for (int ti = 0; ti < p.taps_per_phase; ++ti) kb_total += p.taps[t.phase * p.taps_per_phase + ti].kblocks;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MB * BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = lo0 + static_cast<uint32_t>(stage) * (kStageBytes >> 4);
          const uint32_t b_lo = a_lo + ((MB * kABytes) >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k);
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
                const uint64_t adesc =
                    (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + static_cast<uint32_t>(mb * (kABytes >> 4) + 2 * k));
                // accumulate turns on after EVERY pixel block has issued its first (overwriting) MMA
                if (PAIR) umma_f16_2sm(d_tmem + mb * BN, adesc, bdesc, idesc, (k > 0) ? 1u : accumulate);
                else umma_f16(d_tmem + mb * BN, adesc, bdesc, idesc, (k > 0) ? 1u : accumulate);
              }
            }
            // frees the smem slot (in both CTAs when PAIR) once these MMAs have read it
            if (PAIR) umma_commit_2sm(&empty_bar[stage]);
            else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue(s)
        if (issuer) {
          if (PAIR) umma_commit_2sm(&tfull_bar[acc]);
          else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  };
  auto role_tma_b = [&]() {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer for the weight tiles (HALO)
      int bstage = 0;
      uint32_t bphase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileCoord t = decode_tile(p, tile);
        // K blocks of source 0 (every tap) interleaved with those of source 1 (fused 1x1 shortcut, centre tap): run_steps
        auto step = [&](const int seg, const int kb) {
          const CUtensorMap* bm = &p.b_map[seg];
          const int ntaps = seg == 0 ? p.taps_per_phase : 1;
          
            for (int ti = 0; ti < ntaps; ++ti) {
              const int slab = (seg == 0 ? p.taps[t.phase * p.taps_per_phase + ti].b_slab : p.seg1_slab) +
                               t.frame * p.b_frame_mul;
              mbar_wait_h(&bempty_bar[bstage], bphase ^ 1u, WFK_HINT_NS(p));
              uint8_t* sbt = smem_b + bstage * kBBytes;
              if (PAIR && WFK_XDBG(p) == 8) {   // timing experiment: no weight traffic at all (results are garbage)
                if (leader) mbar_arrive(&bfull_bar[bstage]);
              } else if (PAIR) {
                if (leader) mbar_arrive_expect_tx(&bfull_bar[bstage], 2 * kBBytes);
                tma_load_3d_2sm(sbt, bm, &bfull_bar[bstage], kb * kBlockK,
                                t.nt * BN + static_cast<int>(cta_rank) * kBRows, slab);
              } else {
                mbar_arrive_expect_tx(&bfull_bar[bstage], kBBytes);
                tma_load_3d(sbt, bm, &bfull_bar[bstage], kb * kBlockK, t.nt * BN, slab);
              }
              if (++bstage == kBStages) {
                bstage = 0;
                bphase ^= 1u;
              }
            }
          };
        run_steps(step);
      }
    }
  };
  auto role_xform = [&]() {
    // -------------------------------------------------------------- halo transform (HALO, warps 3..6)
    // GroupNorm apply + SiLU of the consumer's input, fused: y = silu(a[c]*x + b[c]) in place on the staged
    // halo (once per element, reused by all nine taps). Pixels outside the image stay the zeros TMA wrote
    // (the convolution pads AFTER the activation). Without a table the halo is passed through unchanged.
    const int xt = threadIdx.x - 32 * kXformWarp0;  // 0..127
    const int lc = xt & 7;                // logical 16-byte chunk (8 channels) this thread owns
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileCoord t = decode_tile(p, tile);
      const int x0 = (t.tx * kBlocksPerTile + static_cast<int>(cta_rank) * MB) * 8 - 1;
      const int y0 = t.ty * 16 - 1;
      // frame of this CTA's next tile: its (scale, shift) rows are prefetched into L1 during this tile's last K block
      const int next_frame = (tile + tile_step < total_tiles) ? decode_tile(p, tile + tile_step).frame : -1;
      // K blocks of source 0 (every tap) interleaved with those of source 1 (fused 1x1 shortcut, centre tap): run_steps
      auto step = [&](const int seg, const int kb) {
        
          mbar_wait_h(&full_bar[stage], phase, WFK_HINT_NS(p));
          if (seg == 0 && (p.gn_table != nullptr || p.gn_stats != nullptr) && WFK_XDBG(p) != 2) {
            float ga[8], gb[8];
            if (p.gn_stats != nullptr) {
              // same arithmetic as wfk_gn_table: mean / rstd per group in double, (a, b) = (rstd*gamma, beta - mean*a)
              const int c0 = kb * kBlockK + lc * 8;
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + c0));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + c0 + 4));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + c0));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + c0 + 4));
              const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              const double* st = p.gn_stats + static_cast<int64_t>(t.frame) * p.gn_groups * 2;
              float mean_f[2], rstd_f[2];
              const int gfirst = c0 >> p.gn_cpg_log2;
              const int ng = (p.gn_cpg_log2 == 2) ? 2 : 1;   // 8 channels span two groups only when cpg == 4
#pragma unroll
              for (int gi = 0; gi < 2; ++gi) {
                if (gi < ng) {
                  const double mean = st[(gfirst + gi) * 2 + 0] * p.gn_inv_count;
                  double var = st[(gfirst + gi) * 2 + 1] * p.gn_inv_count - mean * mean;
                  var = var < 0.0 ? 0.0 : var;
                  mean_f[gi] = static_cast<float>(mean);
                  rstd_f[gi] = static_cast<float>(rsqrt(var + static_cast<double>(p.gn_eps)));
                } else {
                  mean_f[gi] = mean_f[0];
                  rstd_f[gi] = rstd_f[0];
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int gi = (p.gn_cpg_log2 == 2) ? (j >> 2) : 0;
                const float a = rstd_f[gi] * gam[j];
                // silu(z) = h + h*tanh(h) with h = z/2: one MUFU op per element instead of two (ex2 + rcp)
                ga[j] = 0.5f * a;
                gb[j] = 0.5f * (bet[j] - mean_f[gi] * a);
              }
            } else {
            const float4* tp = reinterpret_cast<const float4*>(p.gn_table + static_cast<int64_t>(t.frame) * p.gn_cin +
                                                               kb * kBlockK + lc * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 v = __ldg(tp + j);
              ga[2 * j] = 0.5f * v.x;
              gb[2 * j] = 0.5f * v.y;
              ga[2 * j + 1] = 0.5f * v.z;
              gb[2 * j + 1] = 0.5f * v.w;
            }
            // The 64 bytes above come from L2 and are consumed at once: every K block began with a full L2 round trip
            // (16 % of the transform warps' samples in the 128-wide kernel). Pull the NEXT K block's rows into L1 now.
            {
              const bool more = kb + 1 < p.seg_kblocks[0];
              const int nf = more ? t.frame : next_frame;
              if (nf >= 0 && (lc & 1) == 0) {
                const float2* np = p.gn_table + static_cast<int64_t>(nf) * p.gn_cin + (more ? kb + 1 : 0) * kBlockK + lc * 8;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
              }
            }
            }
            // Software-pipelined: each thread owns rows (xt>>3) + 16*i of the box; the loads of kXfGroup rows are
            // issued back to back (explicit ld.shared: independent of the stores of the previous group), then
            // transformed and stored. A serial load -> math -> store loop left the transform warps latency-bound
            // and made THEM, not the tensor pipe, the bottleneck of the fused-GroupNorm layers.
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const int row0 = xt >> 3;
#pragma unroll 1
            for (int g0 = 0; g0 < kXfPasses; g0 += kXfGroup) {   // the last group may be partial: `ok` bounds the rows
              uint4 u[kXfGroup];
              uint32_t addr[kXfGroup];
              bool ok[kXfGroup];
#pragma unroll
              for (int j = 0; j < kXfGroup; ++j) {
                const int row = row0 + kXfRowsPerPass * (g0 + j);
                const int hy = row / kHaloPitch, hx = row - hy * kHaloPitch;
                const int py = y0 + hy, px = x0 + hx;
                // zero padding stays zero: out-of-image pixels are skipped
                ok[j] = (g0 + j < kXfPasses) && (row < kHaloPitch * kHaloRows) && py >= 0 && py < p.tile_h && px >= 0 && px < p.tile_w;
                addr[j] = sa + static_cast<uint32_t>(row * 128 + ((lc ^ (row & 7)) << 4));
                if (ok[j]) u[j] = lds_v4(addr[j]);
              }
#pragma unroll
              for (int j = 0; j < kXfGroup; ++j) {
                if (!ok[j]) continue;
                uint32_t* h2 = reinterpret_cast<uint32_t*>(&u[j]);
                if (WFK_XDBG(p) != 1) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    // packed fp32x2 FMAs (sm_100): same rounding per component, half the issue slots
                    const float2 v = __ffma2_rn(A16<BF16>::unpack(h2[e]), make_float2(ga[2 * e], ga[2 * e + 1]),
                                                make_float2(gb[2 * e], gb[2 * e + 1]));
                    float2 t;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(v.x));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(v.y));
                    const float2 y = __ffma2_rn(v, t, v);
                    h2[e] = A16<BF16>::pack(y.x, y.y);
                  }
                }
                sts_v4(addr[j], u[j]);
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to tcgen05.mma
          }
          __syncwarp();
          if (lane == 0) {
            // the writes were made visible to the async proxy by the fence above; a plain (release.cta) remote
            // arrive suffices -- release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR and cost the transform warps
            // a third of their time
            if (PAIR) mbar_arrive_leader_relaxed(&ready_bar[stage]);
            else mbar_arrive(&ready_bar[stage]);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        };
      run_steps(step);
    }
  };
  auto role_epi = [&]() {
    // -------------------------------------------------------------- epilogue (8 warps)
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 32 * kFirstEpiWarp;  // 0..kEpiThreads-1
    const int ew = warp - kFirstEpiWarp;              // epilogue warp index
    const int part = ew >> 2;         // which half of the chunk list this warp owns
    constexpr int kParts = kEpiWarps / 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool do_stats = p.stats != nullptr;
    const bool has_res = p.residual != nullptr;
    int bias_nt = -1;
    if (do_stats) {
      for (int i = et; i < kEpiWarps * (BN / 2); i += kEpiThreads) s_stats[i] = 0.f;
      named_bar_sync(1, kEpiThreads);
    }
    // fixed-order fold of the epilogue warps' partials in s_stats, then one fp64 reduction per (group, moment)
    auto fold_stats = [&](const int frame, const int nt) {
      named_bar_sync(1, kEpiThreads);
      const int groups_in_tile = BN >> p.cpg_log2;
      if (et < 2 * groups_in_tile) {
        const int g = ((nt * BN) >> p.cpg_log2) + (et >> 1);
        float tot = 0.f;
#pragma unroll
        for (int wi = 0; wi < kEpiWarps; ++wi) {
          tot += s_stats[wi * (BN / 2) + et];
          s_stats[wi * (BN / 2) + et] = 0.f;
        }
        if (g < p.groups_total)   // N tail (n_total < BN): groups beyond the tensor only ever held zeros
          atomicAdd(&p.stats[(static_cast<int64_t>(frame) * p.groups_total + g) * 2 + (et & 1)], static_cast<double>(tot));
      }
      named_bar_sync(1, kEpiThreads);
    };
    // Cout = 128 kernel (kRegSplit, 4-channel groups): the (sum, sum of squares) partials of a thread's two chunk
    // columns live in registers across all tiles of a frame; the cross-lane reduction, the shared-memory fold and its
    // two block barriers run once per FRAME CHANGE (~4 tiles) instead of a shuffle butterfly per chunk and a fold per
    // tile -- the epilogue, not the tensor pipe, bounded these K = 9*128 layers.
    const bool reg_stats = kRegSplit && do_stats && p.cpg_log2 == 2 && p.tiles_n == 1;
    float rs0[8], rq0[8], rs1[8], rq1[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) rs0[g] = rq0[g] = rs1[g] = rq1[g] = 0.f;
    int stats_frame = -1;
    auto flush_reg_stats = [&](const int frame) {
      float vals[32];  // index = (slot * 8 + group) * 2 + moment
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        vals[2 * g] = rs0[g];
        vals[2 * g + 1] = rq0[g];
        vals[16 + 2 * g] = rs1[g];
        vals[16 + 2 * g + 1] = rq1[g];
        rs0[g] = rq0[g] = rs1[g] = rq1[g] = 0.f;
      }
      const int k = warp_transpose_reduce32(vals, lane);
      const int slot = k >> 4, g = (k >> 1) & 7, m = k & 1;
      s_stats[ew * (BN / 2) + 2 * (8 * (part + kParts * slot) + g) + m] = vals[0];
      fold_stats(frame, 0);
    };
    // swizzled staging offsets: this lane's OWN row (unit j) and the coalesced (row 8j + lane/4, unit lane & 3) slots
    const uint32_t stg = smem_u32(s_stage) + static_cast<uint32_t>(ew) * kStageWarpBytes;
    const uint32_t own_row = stg + static_cast<uint32_t>(lane & (kStRows - 1)) * 64u;
    const uint32_t own_x = static_cast<uint32_t>((lane >> 1) & 3);
    const int own_pass = lane / kStRows;           // the pass in which this lane's own row sits in the staging tile
    uint32_t co_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t rr = 8u * j + (lane >> 2);
      co_off[j] = stg + rr * 64u + ((static_cast<uint32_t>(lane & 3) ^ ((rr >> 1) & 3u)) << 4);
    }
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileCoord t = decode_tile(p, tile);
      const int row = quarter * 32 + lane;
      if (reg_stats && t.frame != stats_frame) {
        if (stats_frame >= 0) flush_reg_stats(stats_frame);
        stats_frame = t.frame;
      }
      // stage this N tile's bias once (smem broadcast reads in the chunk loop). Without any per-channel vector (attention
      // scores: nine N tiles per pixel block, i.e. a new N tile -- and two barriers over the 256 epilogue threads -- on
      // EVERY tile) the zeros / ones staged for the first tile stay valid.
      if (t.nt != bias_nt && (bias_nt < 0 || p.bias != nullptr || (EPI && (p.scale2 != nullptr || p.shift2 != nullptr)))) {
        named_bar_sync(2, kEpiThreads);
        for (int i = et; i < BN; i += kEpiThreads) {
          const int c = t.nt * BN + i;
          s_bias[i] = (p.bias != nullptr && c < p.n_total) ? __ldg(p.bias + c) : 0.f;
          if constexpr (EPI) {
            s_sc2[i] = (p.scale2 != nullptr && c < p.n_total) ? __ldg(p.scale2 + c) : 1.f;
            s_sh2[i] = (p.shift2 != nullptr && c < p.n_total) ? __ldg(p.shift2 + c) : 0.f;
          }
        }
        named_bar_sync(2, kEpiThreads);
        bias_nt = t.nt;
      }
      // output coordinates of this thread's row in each of this CTA's MB pixel blocks
      bool valid_mb[2] = {false, false};
      int64_t base_mb[2] = {0, 0};
      [[maybe_unused]] int64_t pix_mb[2] = {0, 0};
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int blk = static_cast<int>(cta_rank) * MB + mb;
        const int bx = p.stack_x ? t.tx * kBlocksPerTile + blk : t.tx;
        const int by = p.stack_x ? t.ty : t.ty * kBlocksPerTile + blk;
        const int py = by * bh + (row >> p.bw_log2);
        const int px = bx * bw + (row & (bw - 1));
        valid_mb[mb] = (py < p.tile_h) && (px < p.tile_w);
        const int oy = py * p.out_sy + (t.phase >> 1);
        const int ox = px * p.out_sx + (t.phase & 1);
        const int64_t pix = (static_cast<int64_t>(t.frame) * p.out_rows + oy) * p.out_cols + ox;
        base_mb[mb] = pix * p.ldc + static_cast<int64_t>(t.nt) * BN;
        pix_mb[mb] = pix;
      }
      // Row-softmax passes: this thread's row operand (ROW_EXP: -max * scale, ROW_NORM: 1 / sum; the slots are folded in
      // a fixed order) and its running max / sum over the tile's columns.
      [[maybe_unused]] float row_a[2] = {0.f, 0.f};
      [[maybe_unused]] float row_r[2] = {0.f, 0.f};
      if constexpr (kRowModes) {
        if (p.act >= WFK_ACT_ROW_MAX) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            row_r[mb] = (p.act == WFK_ACT_ROW_MAX) ? -INFINITY : 0.f;
            if (p.act != WFK_ACT_ROW_MAX && valid_mb[mb]) {
              const float* rp = p.row_in + pix_mb[mb] * p.row_ld;
              float a = (p.act == WFK_ACT_ROW_EXP) ? -INFINITY : 0.f;
              for (int i = 0; i < p.row_ld; ++i) {
                const float x = __ldg(rp + i);
                a = (p.act == WFK_ACT_ROW_EXP) ? fmaxf(a, x) : a + x;
              }
              row_a[mb] = (p.act == WFK_ACT_ROW_EXP) ? -a * p.row_scale : 1.f / a;
            }
          }
        }
      }
      const int ncols = min(BN, p.n_total - t.nt * BN);  // N tail: columns >= ncols are padding
      const int nchunks = (ncols + 31) >> 5;
      const int total_it = MB * nchunks;
      // Coalesced global accesses (fp16 output, fp16 residual): for access j a lane serves 16 bytes (unit lane & 3)
      // of row 8*j + lane/4 of the warp's 32 rows, so one instruction covers 8 rows x 64 contiguous bytes instead of
      // 32 rows x 16 bytes. A warp-wide 128-bit store that touches 32 different lines made the epilogue of the
      // K = 9*128 layers LSU-bound (measured: +23 % with the stores removed, +30 % on residual layers).
      int64_t cbase[MB][4];
      uint32_t cvalid = 0;
      const bool tma_store = kTmaStore && p.use_tma_store != 0;
      if (kCoalesce && (!tma_store || has_res)) {   // per-lane output rows: the store path without TMA, the residual loads
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const int blk = static_cast<int>(cta_rank) * MB + mb;
          const int bx = p.stack_x ? t.tx * kBlocksPerTile + blk : t.tx;
          const int by = p.stack_x ? t.ty : t.ty * kBlocksPerTile + blk;
          if (p.bw_log2 == 3) {
            // 8-pixel-wide blocks (every HALO tile): access j is simply j image rows below access 0
            const int rr = quarter * 32 + (lane >> 2);
            const int py = by * bh + (rr >> 3);
            const int px = bx * 8 + (rr & 7);
            const int oy = py * p.out_sy + (t.phase >> 1);
            const int ox = px * p.out_sx + (t.phase & 1);
            const int64_t pix = (static_cast<int64_t>(t.frame) * p.out_rows + oy) * p.out_cols + ox;
            const int64_t b0 = pix * p.ldc + static_cast<int64_t>(t.nt) * BN + 8 * (lane & 3);
            const int64_t row_step = static_cast<int64_t>(p.out_sy) * p.out_cols * p.ldc;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              cbase[mb][j] = b0 + j * row_step;
              if (py + j < p.tile_h && px < p.tile_w) cvalid |= 1u << (mb * 4 + j);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int rr = quarter * 32 + 8 * j + (lane >> 2);
              const int py = by * bh + (rr >> p.bw_log2);
              const int px = bx * bw + (rr & (bw - 1));
              if (py < p.tile_h && px < p.tile_w) cvalid |= 1u << (mb * 4 + j);
              const int oy = py * p.out_sy + (t.phase >> 1);
              const int ox = px * p.out_sx + (t.phase & 1);
              const int64_t pix = (static_cast<int64_t>(t.frame) * p.out_rows + oy) * p.out_cols + ox;
              cbase[mb][j] = pix * p.ldc + static_cast<int64_t>(t.nt) * BN + 8 * (lane & 3);
            }
          }
        }
      }
      // fp32 output rows for the coalesced store: access i serves 16 bytes (unit lane & 7) of row 4*i + lane/8
      int64_t fbase[kCoalesceF32 ? MB : 1][8];
      uint32_t fvalid = 0;
      if constexpr (kCoalesceF32) {
        if (p.out_f != nullptr) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const int blk = static_cast<int>(cta_rank) * MB + mb;
            const int bx = p.stack_x ? t.tx * kBlocksPerTile + blk : t.tx;
            const int by = p.stack_x ? t.ty : t.ty * kBlocksPerTile + blk;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = quarter * 32 + 4 * i + (lane >> 3);
              const int py = by * bh + (rr >> p.bw_log2);
              const int px = bx * bw + (rr & (bw - 1));
              if (py < p.tile_h && px < p.tile_w) fvalid |= 1u << (mb * 8 + i);
              const int oy = py * p.out_sy + (t.phase >> 1);
              const int ox = px * p.out_sx + (t.phase & 1);
              const int64_t pix = (static_cast<int64_t>(t.frame) * p.out_rows + oy) * p.out_cols + ox;
              fbase[mb][i] = pix * p.ldc + static_cast<int64_t>(t.nt) * BN + 4 * (lane & 7);
            }
          }
        }
      }
      uint4 rnext[4];
      auto issue_residual = [&](int it) {
        const int mb_i = (MB == 1) ? 0 : (it >= nchunks ? 1 : 0);
        const int c0_i = (it - mb_i * nchunks) << 5;
        if constexpr (kCoalesce) {
          if (has_res && c0_i + 8 * (lane & 3) < ncols) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int64_t b_i = (MB == 1) ? cbase[0][j] : (mb_i ? cbase[MB - 1][j] : cbase[0][j]);
              if (cvalid & (1u << (mb_i * 4 + j))) rnext[j] = __ldg(reinterpret_cast<const uint4*>(p.residual + b_i + c0_i));
            }
          }
        } else {
          const bool v_i = (MB == 1) ? valid_mb[0] : (mb_i ? valid_mb[1] : valid_mb[0]);
          const int64_t b_i = (MB == 1) ? base_mb[0] : (mb_i ? base_mb[1] : base_mb[0]);
          if (has_res && v_i) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + b_i + c0_i);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c0_i + 8 * j < ncols) rnext[j] = __ldg(rp + j);
          }
        }
      };
      if (part < total_it) issue_residual(part);
      // The residual tensor was written a layer or two ago and has long left L2: pull this tile's rows into L2 now,
      // a whole MMA phase before they are needed, so that the chunk-ahead register prefetch above only has to cover
      // an L2 hit instead of an HBM round trip (which made the "+residual" layers epilogue-bound).
      if (has_res) {
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          if ((MB == 1 || (mb & (kParts - 1)) == part) && valid_mb[mb]) {
            const char* rp = reinterpret_cast<const char*>(p.residual + base_mb[mb]);
            for (int off = (MB == 1 ? part * 128 : 0); off < ncols * 2; off += (MB == 1 ? kParts * 128 : 128))
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + off));
          }
        }
      }
      mbar_wait_h(&tfull_bar[acc], acc_phase, WFK_HINT_NS(p));
      tc_fence_after();
      const uint32_t tlane =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * MB * BN);
#pragma unroll 1
      for (int it = part; it < total_it; it += kParts) {
        const int mb = (MB == 1) ? 0 : (it >= nchunks ? 1 : 0);
        const int c0 = (it - mb * nchunks) << 5;
        const bool valid = (MB == 1) ? valid_mb[0] : (mb ? valid_mb[1] : valid_mb[0]);
        const int64_t base = (MB == 1) ? base_mb[0] : (mb ? base_mb[1] : base_mb[0]);
        uint32_t r[32];
        tmem_ld_32x32b_x32(tlane + static_cast<uint32_t>(mb * BN + c0), r);
        uint4 rcur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
        if (it + kParts < total_it) issue_residual(it + kParts);
        tmem_ld_wait();
        const int nvec = min(4, (ncols - c0) >> 3);  // valid 8-channel vectors in this chunk
        float v[32];
        {
          const uint32_t b4 = smem_u32(s_bias + c0);   // explicit ld.shared (a generic LD through the pointer is slower)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 bu = lds_v4(b4 + 16 * j);
            const float4 b = make_float4(__uint_as_float(bu.x), __uint_as_float(bu.y), __uint_as_float(bu.z), __uint_as_float(bu.w));
            const float2 s01 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])),
                                          make_float2(b.x, b.y));
            const float2 s23 = __fadd2_rn(make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])),
                                          make_float2(b.z, b.w));
            v[4 * j + 0] = s01.x;
            v[4 * j + 1] = s01.y;
            v[4 * j + 2] = s23.x;
            v[4 * j + 3] = s23.y;
          }
        }
        if (has_res) {
          uint4 rrow[4];
          if constexpr (kCoalesce) {
            if (tma_store) {   // the staging tile may still be the source of the previous chunk's TMA store
              if (lane == 0) bulk_wait_group_read<0>();
              __syncwarp();
            }
            // coalesced pieces -> staging -> this lane's own row
#pragma unroll
            for (int ps = 0; ps < kStPasses; ++ps) {
#pragma unroll
              for (int j = 0; j < kStAcc; ++j) sts_v4(co_off[j], rcur[ps * kStAcc + j]);
              __syncwarp();
              if (kStPasses == 1 || own_pass == ps) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rrow[j] = lds_v4(own_row + ((static_cast<uint32_t>(j) ^ own_x) << 4));
              }
              __syncwarp();
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rrow[j] = rcur[j];
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < nvec) {
                const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&rrow[j]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = A16<BF16>::unpack(h2[e]);
                  v[8 * j + 2 * e + 0] += f.x;
                  v[8 * j + 2 * e + 1] += f.y;
                }
              }
            }
          }
        }
        if constexpr (EPI) {
          if (kRowModes && p.act >= WFK_ACT_ROW_MAX) {
            const int nv = ncols - c0;   // valid columns of this chunk (the accumulators beyond hold zero-filled padding)
            const float ra = (MB == 1) ? row_a[0] : (mb ? row_a[1] : row_a[0]);
            float rr = (MB == 1) ? row_r[0] : (mb ? row_r[1] : row_r[0]);
            if (p.act == WFK_ACT_ROW_MAX) {
#pragma unroll
              for (int i = 0; i < 32; ++i) rr = fmaxf(rr, i < nv ? v[i] : -INFINITY);
            } else if (p.act == WFK_ACT_ROW_EXP) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float e = i < nv ? ex2_approx(fmaf(v[i], p.row_scale, ra)) : 0.f;
                v[i] = e;
                rr += e;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], ra, s_sh2[c0 + i]);
            }
            if (MB == 1 || mb == 0) row_r[0] = rr;
            else row_r[1] = rr;
          } else if (p.act != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = act_apply(p.act, v[i], p.act_slope);
          }
        }
        if (reg_stats) {
          auto accum = [&](float (&rs)[8], float (&rq)[8]) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                a += v[4 * g + j];
                b = fmaf(v[4 * g + j], v[4 * g + j], b);
              }
              rs[g] += valid ? a : 0.f;
              rq[g] += valid ? b : 0.f;
            }
          };
          if ((c0 >> 5) < kParts) accum(rs0, rq0);   // chunk column part           (slot 0)
          else accum(rs1, rq1);                      // chunk column part + kParts  (slot 1)
        } else if (do_stats) {
          float* dst = s_stats + ew * (BN / 2) + 2 * (c0 >> p.cpg_log2);
          if (p.cpg_log2 == 2) chunk_group_stats<8>(v, valid, lane, dst);
          else if (p.cpg_log2 == 3) chunk_group_stats<4>(v, valid, lane, dst);
          else chunk_group_stats<2>(v, valid, lane, dst);
        }
        if (!kCoalesce) {
          if (valid && p.out_h != nullptr && WFK_XDBG(p) != 32) {
            uint4* op = reinterpret_cast<uint4*>(p.out_h + base + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < nvec) {
                uint4 u;
                uint32_t* h2 = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = A16<BF16>::pack(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                op[j] = u;
              }
            }
          }
        } else if (p.out_h != nullptr && WFK_XDBG(p) != 32) {   // 32: timing experiment without the fp16 stores
          uint4 up[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t* h2 = reinterpret_cast<uint32_t*>(&up[j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = A16<BF16>::pack(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
          }
          if (tma_store) {
#pragma unroll
            for (int ps = 0; ps < kStPasses; ++ps) {
              // the previous store must have finished READING the staging tile (issued one chunk / pass of work ago)
              if (lane == 0) bulk_wait_group_read<0>();
              __syncwarp();
              if (kStPasses == 1 || own_pass == ps) {
#pragma unroll
                for (int j = 0; j < 4; ++j) sts_v4(own_row + ((static_cast<uint32_t>(j) ^ own_x) << 4), up[j]);
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the TMA engine
              __syncwarp();
              if (lane == 0) {
                const int blk = static_cast<int>(cta_rank) * MB + mb;
                const int sx = (t.tx * kBlocksPerTile + blk) * 8, sy = t.ty * 16 + quarter * 4 + ps * (kStRows / 8);
                if (p.use_tma_store == 2)   // sub-pixel phase of the fused upsample: (c, x parity, x, y parity, frame * h + y)
                  tma_store_5d(&p.out_map, stg, t.nt * BN + c0, t.phase & 1, sx, t.phase >> 1, t.frame * p.tile_h + sy);
                else
                  tma_store_4d(&p.out_map, stg, t.nt * BN + c0, sx, sy, t.frame);
                bulk_commit_group();
              }
            }
          } else
#pragma unroll
          for (int ps = 0; ps < kStPasses; ++ps) {
            if (kStPasses == 1 || own_pass == ps) {
#pragma unroll
              for (int j = 0; j < 4; ++j) sts_v4(own_row + ((static_cast<uint32_t>(j) ^ own_x) << 4), up[j]);
            }
            __syncwarp();
            if (c0 + 8 * (lane & 3) < ncols) {
#pragma unroll
              for (int j = 0; j < kStAcc; ++j) {
                const int jj = ps * kStAcc + j;
                const int64_t b_j = (MB == 1) ? cbase[0][jj] : (mb ? cbase[MB - 1][jj] : cbase[0][jj]);
                if (cvalid & (1u << (mb * 4 + jj))) *reinterpret_cast<uint4*>(p.out_h + b_j + c0) = lds_v4(co_off[j]);
              }
            }
            __syncwarp();
          }
        }
        if constexpr (kCoalesceF32) {
          if (p.out_f != nullptr) {
            // own row (32 floats = 8 units of 16 B, unit XOR-swizzled by the row) -> staging -> 4 rows x 128 B per store
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_v4(stg + static_cast<uint32_t>(lane) * 128u + ((static_cast<uint32_t>(j) ^ (lane & 7u)) << 4),
                     make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                __float_as_uint(v[4 * j + 3])));
            __syncwarp();
            if (c0 + 4 * (lane & 7) < ncols) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t rr = 4u * i + (lane >> 3);
                const int64_t b_i = (MB == 1) ? fbase[0][i] : (mb ? fbase[MB - 1][i] : fbase[0][i]);
                if (fvalid & (1u << (mb * 8 + i)))
                  *reinterpret_cast<uint4*>(p.out_f + b_i + c0) =
                      lds_v4(stg + rr * 128u + (((lane & 7u) ^ (rr & 7u)) << 4));
              }
            }
            __syncwarp();
          }
        }
        if (valid) {
          if (!kCoalesceF32 && p.out_f != nullptr) {
            float4* op = reinterpret_cast<float4*>(p.out_f + base + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < 2 * nvec) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if constexpr (EPI) {
            if (p.out2_h != nullptr) {
              uint4* op = reinterpret_cast<uint4*>(p.out2_h + base + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j < nvec) {
                  uint4 u;
                  uint32_t* h2 = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int i0 = 8 * j + 2 * e;
                    const float a = act_apply(p.act2, fmaf(v[i0], s_sc2[c0 + i0], s_sh2[c0 + i0]), p.act_slope);
                    const float b = act_apply(p.act2, fmaf(v[i0 + 1], s_sc2[c0 + i0 + 1], s_sh2[c0 + i0 + 1]), p.act_slope);
                    h2[e] = A16<BF16>::pack(a, b);
                  }
                  op[j] = u;
                }
              }
            }
          }
        }
      }  // chunk loop
      if constexpr (kRowModes) {
        if (p.act >= WFK_ACT_ROW_MAX && p.row_out != nullptr) {   // one (row, slot) per thread and block
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
            if (valid_mb[mb]) p.row_out[pix_mb[mb] * p.row_ld + t.nt * kParts + part] = row_r[mb];
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp (of the leader CTA when PAIR)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader_relaxed(&tempty_bar[acc]);  // tcgen05.wait::ld + fence::before above order the TMEM reads
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;

      if (do_stats && !reg_stats) fold_stats(t.frame, t.nt);
    }
    if (reg_stats && stats_frame >= 0) flush_reg_stats(stats_frame);
    if (kTmaStore && lane == 0) bulk_wait_group<0>();   // this warp's output stores are complete before the CTA retires
  };

  if constexpr (kRegSplit) {
    // whole warpgroups per role; each group first trades registers (the launch allocation is 96 per thread)
    if (warp >= kWarpTmaB) {
      setmaxnreg_dec<56>();
      if (warp == kWarpTmaA) role_tma_a();
      else if (warp == kWarpMma) role_mma();
      else if (warp == kWarpTmaB) role_tma_b();
    } else if (warp < kXformWarp0 + kXformWarps) {
      setmaxnreg_dec<64>();
      role_xform();
    } else {
      setmaxnreg_inc<144>();
      role_epi();
    }
  } else {
    if (warp == kWarpTmaA) role_tma_a();
    else if (warp == kWarpMma) role_mma();
    else if (HALO && warp == kWarpTmaB) role_tma_b();
    else if (HALO && warp >= kXformWarp0 && warp < kXformWarp0 + kXformWarps) role_xform();
    else role_epi();
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its smem
  else __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm<2 * MB * BN>(tmem_base);
    else tmem_dealloc<2 * MB * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------- host
template <int BN, bool PAIR, bool HALO>
struct ConvCfg;
template <>
struct ConvCfg<256, false, false> {
  static constexpr int kMB = 1, kStages = 3;   // 48 KB / stage (+ 32 KB epilogue staging)
};
template <>
struct ConvCfg<128, false, false> {            // two pixel blocks share the weight tile
  static constexpr int kMB = 2, kStages = 3;   // 48 KB / stage (+ 32 KB epilogue staging)
};
template <>
struct ConvCfg<256, true, false> {
  static constexpr int kMB = 1, kStages = 5;   // 16 + 16 KB / stage (+ 32 KB epilogue staging)
};
template <>
struct ConvCfg<128, true, false> {
  static constexpr int kMB = 2, kStages = 4;   // 32 + 8 KB / stage (+ 32 KB epilogue staging)
};
template <>
struct ConvCfg<256, true, true> {
  static constexpr int kMB = 1, kStages = 5;   // 5 x 23 KB halo boxes + 6 x 16 KB weight tiles
};
template <>
struct ConvCfg<128, true, true> {
  static constexpr int kMB = 2, kStages = 4;   // 4 x 41 KB halo boxes + 5 x 8 KB weight tiles
};

template <int BN, bool PAIR, bool HALO, bool EPI = false>
constexpr size_t conv_smem_bytes() {
  using Cfg = ConvCfg<BN, PAIR, HALO>;
  constexpr size_t b_bytes = static_cast<size_t>(PAIR ? BN / 2 : BN) * kBlockK * 2;
  constexpr size_t halo_stage = (static_cast<size_t>(8 * Cfg::kMB + 2) * kHaloRows * 128 + 1023) & ~static_cast<size_t>(1023);
  constexpr size_t ring = HALO ? Cfg::kStages * halo_stage + (Cfg::kMB == 1 ? (EPI ? 5 : 6) : 5) * b_bytes
                               : Cfg::kStages * (Cfg::kMB * kABytes + b_bytes);
  return 1024 /*align slack*/ + ring + (3 * Cfg::kStages + 2 * kHaloBStagesMax + 4) * 8 + 16 +
         kEpiWarps * (BN / 2) * 4 + BN * 4 * (EPI ? 3 : 1) + kEpiWarps * (HALO ? (BN == 256 ? 1024 : 2048) : 4096) + 64 +
         ((HALO && !EPI) ? 1024 : 0) /* 1 KB alignment of the TMA-store staging tiles */;
}

template <int BN, bool PAIR, bool HALO, bool EPI = false, bool BF16 = false>
cudaError_t launch_conv(const ConvKernelParams& params, int grid, cudaStream_t s) {
  using Cfg = ConvCfg<BN, PAIR, HALO>;
  auto kern = conv_gemm_kernel<BN, Cfg::kMB, Cfg::kStages, PAIR, HALO, EPI, BF16>;
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(conv_smem_bytes<BN, PAIR, HALO, EPI>()));
    if (e != cudaSuccess) return e;
    attr_lock.finished();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(conv_threads(HALO, Cfg::kMB));
  cfg.dynamicSmemBytes = conv_smem_bytes<BN, PAIR, HALO, EPI>();
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = !(std::getenv("WFK_PDL") && std::getenv("WFK_PDL")[0] == '0');
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, params);
}

// Largest number of co-resident 2-CTA clusters for the pair kernels (a persistent grid must not exceed it,
// or the surplus clusters run as a second wave). Queried once per variant.
template <int BN, bool HALO>
int max_active_pairs() {
  using Cfg = ConvCfg<BN, true, HALO>;
  static std::atomic<int> cached_dev[kMaxDevices];   // zero-initialised: 0 = not queried yet (per device)
  const int dev = t_device < 0 ? 0 : t_device;
  if (cached_dev[dev].load() > 0) return cached_dev[dev].load();
  auto kern = conv_gemm_kernel<BN, Cfg::kMB, Cfg::kStages, true, HALO, false, false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       static_cast<int>(conv_smem_bytes<BN, true, HALO>()));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms());
  cfg.blockDim = dim3(conv_threads(HALO, Cfg::kMB));
  cfg.dynamicSmemBytes = conv_smem_bytes<BN, true, HALO>();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = num_sms() / 2;
  if (std::getenv("WFK_DEBUG")) fprintf(stderr, "[wfk] conv_gemm<%d,pair,halo=%d>: max active clusters = %d\n", BN, (int)HALO, n);
  cached_dev[dev].store(n);
  return n;
}

}  // namespace wfk

struct wfk_conv_plan {
  wfk::ConvKernelParams params;
  int bn;
  int pair;
  int halo;
  int epi;   // extended epilogue (activation / second activated output): EPI kernel variants
  int bf16;  // bf16 operands / activations (CTA-pair kernels without the extended epilogue only)
  int grid;
  int device;
};

namespace {

int encode_a(const wfk_view5& v, int bw, int bh, bool bf16, CUtensorMap* out) {
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5] = {static_cast<cuuint32_t>(wfk::kBlockK), static_cast<cuuint32_t>(bw), 1u,
                       static_cast<cuuint32_t>(bh), 1u};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) dims[i] = static_cast<cuuint64_t>(v.dim[i]);
  for (int i = 1; i < 5; ++i) strides[i - 1] = static_cast<cuuint64_t>(v.stride[i]);
  CUresult r = wfk::g_encode_tiled(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                                   const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return wfk::fail(WFK_ERR_CUDA,
                     "cuTensorMapEncodeTiled(A) failed: %d dims=(%lld,%lld,%lld,%lld,%lld) strides=(%lld,%lld,%lld,%lld) "
                     "box=(64,%d,1,%d,1)",
                     (int)r, (long long)v.dim[0], (long long)v.dim[1], (long long)v.dim[2], (long long)v.dim[3],
                     (long long)v.dim[4], (long long)v.stride[1], (long long)v.stride[2], (long long)v.stride[3],
                     (long long)v.stride[4], bw, bh);
  return WFK_OK;
}

int encode_b(const wfk_view3& v, int rows, bool bf16, CUtensorMap* out) {
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3] = {static_cast<cuuint32_t>(wfk::kBlockK), static_cast<cuuint32_t>(rows), 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < 3; ++i) dims[i] = static_cast<cuuint64_t>(v.dim[i]);
  for (int i = 1; i < 3; ++i) strides[i - 1] = static_cast<cuuint64_t>(v.stride[i]);
  CUresult r = wfk::g_encode_tiled(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                                   const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return wfk::fail(WFK_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d dims=(%lld,%lld,%lld) strides=(%lld,%lld)",
                     (int)r, (long long)v.dim[0], (long long)v.dim[1], (long long)v.dim[2], (long long)v.stride[1],
                     (long long)v.stride[2]);
  return WFK_OK;
}

int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

bool halo_mode_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("WFK_CONV_HALO");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return mode == 1;
}

// HALO applies to stride-1 taps within one pixel of the centre on a plain NHWC view (3x3 convolutions, the
// 2x2 sub-pixel phases of the upsampler), optionally followed by ONE centre tap on source 1 (fused shortcut).
bool halo_eligible(const wfk_conv_desc* d, int* n_src0_taps, int* kb0, int* kb1, int* slab1) {
  if (d->tile_h < 2 || d->a[0].dim[2] != 1 || d->a_frame_mul != 1) return false;
  if (d->a[0].dim[1] != d->tile_w || d->a[0].dim[3] != d->tile_h) return false;
  const int tpp = d->taps_per_phase;
  *kb1 = 0;
  *slab1 = 0;
  int n0 = tpp;
  const wfk_tap& last = d->taps[tpp - 1];
  if (last.src == 1) {
    if (d->num_phases != 1 || last.dx != 0 || last.dy != 0 || last.q != 0 || last.c_off != 0) return false;
    if (d->a[1].dim[2] != 1 || d->a[1].dim[1] != d->tile_w || d->a[1].dim[3] != d->tile_h) return false;
    *kb1 = last.kblocks;
    *slab1 = last.b_slab;
    n0 = tpp - 1;
  }
  if (n0 < 1) return false;
  *kb0 = d->taps[0].kblocks;
  for (int ph = 0; ph < d->num_phases; ++ph)
    for (int i = 0; i < n0; ++i) {
      const wfk_tap& t = d->taps[ph * tpp + i];
      if (t.src != 0 || t.q != 0 || t.c_off != 0 || t.dx < -1 || t.dx > 1 || t.dy < -1 || t.dy > 1 || t.kblocks != *kb0)
        return false;
    }
  *n_src0_taps = n0;
  return true;
}

bool pair_mode_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("WFK_CONV_PAIR");
    mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return mode == 1;
}

}  // namespace

extern "C" int wfk_conv_plan_create(const wfk_conv_desc* d, wfk_conv_plan** out) {
  WFK_REQUIRE(d != nullptr && out != nullptr, "null argument");
  WFK_ENTER_PTR(d->a[0].ptr);   // the plan belongs to the device that holds its operands
  WFK_REQUIRE(d->n_total > 0 && d->n_total % 8 == 0, "n_total=%d must be a positive multiple of 8", d->n_total);
  WFK_REQUIRE(d->num_phases == 1 || d->num_phases == 4, "num_phases must be 1 or 4");
  WFK_REQUIRE(d->taps_per_phase >= 1 && d->num_phases * d->taps_per_phase <= WFK_MAX_TAPS, "too many taps");
  WFK_REQUIRE(d->n_frames >= 1 && d->tile_h >= 1 && d->tile_w >= 1, "empty problem");
  WFK_REQUIRE(d->a[0].ptr != nullptr && d->b[0].ptr != nullptr, "A/B source 0 missing");
  WFK_REQUIRE(d->out_h != nullptr || d->out_f != nullptr || d->out2_h != nullptr ||
                  (d->act == WFK_ACT_ROW_MAX && d->row_out != nullptr),
              "no output requested");
  WFK_REQUIRE(d->act >= 0 && d->act <= WFK_ACT_ROW_NORM && d->act2 >= 0 && d->act2 <= WFK_ACT_SILU, "unknown activation");
  if (d->act >= WFK_ACT_ROW_MAX) {
    WFK_REQUIRE(d->act == WFK_ACT_ROW_NORM || d->row_out != nullptr, "ROW_MAX / ROW_EXP need row_out");
    WFK_REQUIRE(d->act == WFK_ACT_ROW_MAX || d->row_in != nullptr, "ROW_EXP / ROW_NORM need row_in");
    WFK_REQUIRE(d->act != WFK_ACT_ROW_EXP || d->out_h != nullptr, "ROW_EXP writes its values to out_h");
    WFK_REQUIRE(d->residual == nullptr && d->stats == nullptr && d->out2_h == nullptr && d->num_phases == 1 &&
                    d->taps_per_phase == 1,
                "row-softmax epilogues take plain one-tap GEMM descriptors");
  }
  WFK_REQUIRE(d->ldc % 8 == 0, "ldc must be a multiple of 8");
  bool uses_src1 = false;
  for (int i = 0; i < d->num_phases * d->taps_per_phase; ++i) {
    const wfk_tap& t = d->taps[i];
    WFK_REQUIRE(t.src == 0 || t.src == 1, "tap %d: bad src", i);
    WFK_REQUIRE(t.kblocks >= 1, "tap %d: kblocks must be >= 1", i);
    WFK_REQUIRE(t.c_off % 8 == 0, "tap %d: c_off must be a multiple of 8", i);
    uses_src1 |= (t.src == 1);
  }
  if (uses_src1) WFK_REQUIRE(d->a[1].ptr != nullptr && d->b[1].ptr != nullptr, "A/B source 1 missing");
  int cpg_log2 = 0;
  if (d->stats != nullptr) {
    cpg_log2 = ilog2_exact(d->cpg);
    WFK_REQUIRE(cpg_log2 >= 2 && cpg_log2 <= 4, "cpg=%d must be 4, 8 or 16 when stats are requested", d->cpg);
  }

  wfk_conv_plan* plan = new (std::nothrow) wfk_conv_plan();
  WFK_REQUIRE(plan != nullptr, "out of memory");
  wfk::ConvKernelParams& p = plan->params;
  plan->bn = (d->n_total % 256 == 0 || d->n_total > 256) ? 256 : 128;
  plan->pair = (pair_mode_enabled() && wfk::num_sms() % 2 == 0) ? 1 : 0;
  if (d->stats != nullptr && d->n_total % 32 != 0) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "stats need n_total (%d) to be a multiple of 32 (one accumulator chunk)", d->n_total);
  }

  // 128-pixel blocks per tile: MB per CTA, x2 for a CTA pair
  const int mb_blocks = (plan->bn == 256 ? 1 : 2) * (plan->pair ? 2 : 1);
  int n_src0_taps = d->taps_per_phase, kb0 = 0, kb1 = 0, slab1 = 0;
  plan->halo = (plan->pair && halo_mode_enabled() && halo_eligible(d, &n_src0_taps, &kb0, &kb1, &slab1)) ? 1 : 0;
  // GEMM-like problems (one row of pixels) stack a tile's blocks along x, images along y
  const int stack_x = (d->tile_h == 1 || plan->halo) ? 1 : 0;
  // block geometry: BW x BH = 128 output pixels, BW a power of two; minimise the padded area
  int best_log2 = 7;
  long best_area = -1;
  for (int l = plan->halo ? 3 : 7; l >= 3; --l) {
    const int bw = 1 << l, bh = 128 >> l;
    const int ew = stack_x ? bw * mb_blocks : bw, eh = stack_x ? bh : bh * mb_blocks;
    const long area = static_cast<long>((d->tile_w + ew - 1) / ew) * ew * (static_cast<long>((d->tile_h + eh - 1) / eh) * eh);
    if (best_area < 0 || area < best_area) {
      best_area = area;
      best_log2 = l;
    }
  }
  const int bw = 1 << best_log2, bh = 128 >> best_log2;
  const int ew = stack_x ? bw * mb_blocks : bw, eh = stack_x ? bh : bh * mb_blocks;
  p.bw_log2 = best_log2;
  p.stack_x = stack_x;
  p.n_frames = d->n_frames;
  p.tile_h = d->tile_h;
  p.tile_w = d->tile_w;
  p.tiles_x = (d->tile_w + ew - 1) / ew;
  p.tiles_y = (d->tile_h + eh - 1) / eh;
  p.tiles_n = (d->n_total + plan->bn - 1) / plan->bn;
  p.n_total = d->n_total;
  p.num_phases = d->num_phases;
  p.taps_per_phase = plan->halo ? n_src0_taps : d->taps_per_phase;
  p.seg_kblocks[0] = kb0;
  p.seg_kblocks[1] = kb1;
  p.seg1_slab = slab1;
  p.gn_table = static_cast<const float2*>(d->gn_table);
  p.gn_cin = static_cast<int>(d->a[0].dim[0]);
  p.gn_stats = d->gn_stats;
  p.gn_gamma = d->gn_gamma;
  p.gn_beta = d->gn_beta;
  p.gn_eps = d->gn_eps;
  p.gn_groups = d->gn_groups;
  p.gn_cpg_log2 = 0;
  p.gn_inv_count = 0.0;
  if (d->gn_stats != nullptr) {
    const int cin = static_cast<int>(d->a[0].dim[0]);
    const int cpg_in = (d->gn_groups > 0 && cin % d->gn_groups == 0) ? cin / d->gn_groups : 0;
    const int l2 = ilog2_exact(cpg_in);
    if (d->gn_gamma == nullptr || d->gn_beta == nullptr || l2 < 2 || d->gn_table != nullptr) {
      delete plan;
      return wfk::fail(WFK_ERR_INVALID, "gn_stats needs gamma, beta, >= 4 (power of two) channels per group and no gn_table");
    }
    p.gn_cpg_log2 = l2;
    p.gn_inv_count = 1.0 / (static_cast<double>(cpg_in) * d->tile_h * d->tile_w);
  }
  p.wait_hint_ns = 0;
  p.xform_debug = 0;
#ifdef WFK_EXPERIMENTS
  p.wait_hint_ns = std::getenv("WFK_WAIT_HINT") ? std::atoi(std::getenv("WFK_WAIT_HINT")) : 0;
  p.xform_debug = std::getenv("WFK_XFORM_DEBUG") ? std::atoi(std::getenv("WFK_XFORM_DEBUG")) : 0;
#endif
  if ((d->gn_table != nullptr || d->gn_stats != nullptr) && !plan->halo) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "a fused GroupNorm+SiLU input (gn_table) needs a HALO-eligible 3x3 stride-1 convolution");
  }
  p.a_frame_mul = d->a_frame_mul;
  p.b_frame_mul = d->b_frame_mul;
  p.bias = d->bias;
  p.residual = static_cast<const __half*>(d->residual);
  p.out_h = static_cast<__half*>(d->out_h);
  p.out_f = d->out_f;
  p.stats = d->stats;
  p.out_rows = d->out_rows;
  p.out_cols = d->out_cols;
  p.out_sy = d->out_sy;
  p.out_sx = d->out_sx;
  p.ldc = d->ldc;
  p.act = d->act;
  p.act2 = d->act2;
  p.act_slope = d->act_slope;
  p.out2_h = static_cast<__half*>(d->out2_h);
  p.scale2 = d->scale2;
  p.shift2 = d->shift2;
  p.row_in = d->row_in;
  p.row_out = d->row_out;
  p.row_ld = d->row_ld;
  p.row_scale = d->row_scale;
  plan->epi = (d->act != 0 || d->out2_h != nullptr) ? 1 : 0;
  plan->bf16 = d->operand_bf16 ? 1 : 0;
  if (plan->bf16 && (plan->epi || !plan->pair)) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "bf16 operands are built for the CTA-pair kernels without activation epilogues only");
  }
  if (plan->epi && !plan->pair) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "activation epilogues need the CTA-pair kernels (WFK_CONV_PAIR=0 is set)");
  }
  // ROW_MAX / ROW_EXP write one slot per (N tile, epilogue half); ROW_NORM only reads the producer's slots
  if (d->act >= WFK_ACT_ROW_MAX &&
      (plan->halo || d->row_ld < 1 || (d->act != WFK_ACT_ROW_NORM && d->row_ld != 2 * p.tiles_n))) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "row-softmax epilogue: row_ld=%d must be 2 * ceil(n_total / %d) = %d (and no 3x3 taps)",
                     d->row_ld, plan->bn, 2 * p.tiles_n);
  }
  p.cpg_log2 = cpg_log2;
  p.groups_total = d->stats ? (d->n_total >> cpg_log2) : 0;
  p.idesc = wfk::umma_idesc_f16(plan->pair ? 256u : 128u, static_cast<uint32_t>(plan->bn), d->operand_bf16 ? 1u : 0u);
  for (int i = 0; i < WFK_MAX_TAPS; ++i) p.taps[i] = d->taps[i];
  if (plan->halo && n_src0_taps != d->taps_per_phase)  // drop the shortcut tap from the (single-phase) tap list
    for (int i = 0; i < n_src0_taps; ++i) p.taps[i] = d->taps[i];

  const int b_rows = plan->pair ? plan->bn / 2 : plan->bn;
  const int mb_cta = plan->bn == 256 ? 1 : 2;
  const int a_bw = plan->halo ? 8 * mb_cta + 2 : bw, a_bh = plan->halo ? wfk::kHaloRows : bh;
  int rc = encode_a(d->a[0], a_bw, a_bh, d->operand_bf16 != 0, &p.a_map[0]);
  if (rc == WFK_OK) rc = encode_b(d->b[0], b_rows, d->operand_bf16 != 0, &p.b_map[0]);
  if (rc == WFK_OK && uses_src1) {
    rc = plan->halo ? encode_a(d->a[1], 8 * mb_cta, 16, d->operand_bf16 != 0, &p.a_map[1])
                    : encode_a(d->a[1], a_bw, a_bh, d->operand_bf16 != 0, &p.a_map[1]);
    if (rc == WFK_OK) rc = encode_b(d->b[1], b_rows, d->operand_bf16 != 0, &p.b_map[1]);
  } else if (rc == WFK_OK) {
    p.a_map[1] = p.a_map[0];
    p.b_map[1] = p.b_map[0];
  }
  if (rc != WFK_OK) {
    delete plan;
    return rc;
  }
  p.use_tma_store = 0;
  static const bool tma_store_enabled = !(std::getenv("WFK_TMA_STORE") && std::getenv("WFK_TMA_STORE")[0] == '0');   // A/B switch
  if (tma_store_enabled && plan->halo && !plan->epi && d->out_h != nullptr && d->out_f == nullptr &&
      d->n_total % plan->bn == 0 && d->num_phases == 1 && d->out_sy == 1 && d->out_sx == 1 && d->out_rows == d->tile_h &&
      d->out_cols == d->tile_w && (reinterpret_cast<uintptr_t>(d->out_h) & 15) == 0) {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->n_total), static_cast<cuuint64_t>(d->out_cols),
                          static_cast<cuuint64_t>(d->out_rows), static_cast<cuuint64_t>(d->n_frames)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->ldc) * 2, static_cast<cuuint64_t>(d->out_cols) * d->ldc * 2,
                             static_cast<cuuint64_t>(d->out_rows) * d->out_cols * d->ldc * 2};
    cuuint32_t box[4] = {32u, 8u, plan->bn == 128 ? 4u : 2u, 1u};   // one staging pass of an epilogue warp
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = wfk::g_encode_tiled(&p.out_map, d->operand_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                     4, d->out_h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) p.use_tma_store = 1;   // otherwise: the per-lane store path
  }
  // The four sub-pixel phases of the fused nearest-x2 upsample write every other pixel of every other row. Seen as
  // (channel, x parity, x, y parity, frame * tile_h + y) the output of ONE phase is a plain 5-D tiled tensor (the frames
  // merge into the row axis because a frame is exactly tile_h row pairs), so the same 32-channel x 8 x 4|2 pixel boxes
  // apply. Whole 16-row tiles only: a ragged last tile would spill into the next frame instead of being clipped.
  static const bool tma_store_up_enabled = !(std::getenv("WFK_TMA_STORE_UP") && std::getenv("WFK_TMA_STORE_UP")[0] == '0');   // A/B switch
  if (tma_store_enabled && tma_store_up_enabled && plan->halo && !plan->epi && d->out_h != nullptr && d->out_f == nullptr &&
      d->n_total % plan->bn == 0 && d->num_phases == 4 && d->out_sy == 2 && d->out_sx == 2 && d->out_rows == 2 * d->tile_h &&
      d->out_cols == 2 * d->tile_w && d->tile_h % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out_h) & 15) == 0) {
    const cuuint64_t px = static_cast<cuuint64_t>(d->ldc) * 2, row = static_cast<cuuint64_t>(d->out_cols) * px;
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(d->n_total), 2, static_cast<cuuint64_t>(d->tile_w), 2,
                          static_cast<cuuint64_t>(d->n_frames) * d->tile_h};
    cuuint64_t strides[4] = {px, 2 * px, row, 2 * row};
    cuuint32_t box[5] = {32u, 1u, 8u, 1u, plan->bn == 128 ? 4u : 2u};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = wfk::g_encode_tiled(&p.out_map, d->operand_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                     5, d->out_h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) p.use_tma_store = 2;
  }
  const long total_tiles = static_cast<long>(p.num_phases) * p.n_frames * p.tiles_y * p.tiles_x * p.tiles_n;
  if (total_tiles > 0x3fffffffL) {
    delete plan;
    return wfk::fail(WFK_ERR_INVALID, "too many tiles");
  }
  if (plan->pair) {
    const long pairs = plan->halo ? (plan->bn == 256 ? wfk::max_active_pairs<256, true>() : wfk::max_active_pairs<128, true>())
                                  : (plan->bn == 256 ? wfk::max_active_pairs<256, false>() : wfk::max_active_pairs<128, false>());
    long use = pairs;
    // experiment knob: run on fewer CTA pairs (what a 4-CTA-cluster design would get: 66 of 74) to see how much of the
    // lost SMs the power cap gives back as clock
#ifdef WFK_EXPERIMENTS
    if (const char* e = std::getenv("WFK_MAX_PAIRS")) {
      const long cap = std::atol(e);
      if (cap >= 1 && cap < use) use = cap;
    }
#endif
    plan->grid = static_cast<int>(2 * (total_tiles < use ? total_tiles : use));
  } else {
    plan->grid = static_cast<int>(total_tiles < wfk::num_sms() ? total_tiles : wfk::num_sms());
  }
  plan->device = wfk::t_device;
  *out = plan;
  return WFK_OK;
}

extern "C" int wfk_conv_plan_run(const wfk_conv_plan* plan, void* stream) {
  WFK_REQUIRE(plan != nullptr, "null plan");
  WFK_ENTER_DEVICE(plan->device);   // the device that holds the plan's operands; `stream` must belong to it
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (plan->epi) {
    if (plan->halo)
      e = (plan->bn == 256) ? wfk::launch_conv<256, true, true, true>(plan->params, plan->grid, s)
                            : wfk::launch_conv<128, true, true, true>(plan->params, plan->grid, s);
    else
      e = (plan->bn == 256) ? wfk::launch_conv<256, true, false, true>(plan->params, plan->grid, s)
                            : wfk::launch_conv<128, true, false, true>(plan->params, plan->grid, s);
  } else if (plan->bf16) {
    if (plan->halo)
      e = (plan->bn == 256) ? wfk::launch_conv<256, true, true, false, true>(plan->params, plan->grid, s)
                            : wfk::launch_conv<128, true, true, false, true>(plan->params, plan->grid, s);
    else
      e = (plan->bn == 256) ? wfk::launch_conv<256, true, false, false, true>(plan->params, plan->grid, s)
                            : wfk::launch_conv<128, true, false, false, true>(plan->params, plan->grid, s);
  } else if (plan->halo) {
    e = (plan->bn == 256) ? wfk::launch_conv<256, true, true>(plan->params, plan->grid, s)
                          : wfk::launch_conv<128, true, true>(plan->params, plan->grid, s);
  } else if (plan->pair) {
    e = (plan->bn == 256) ? wfk::launch_conv<256, true, false>(plan->params, plan->grid, s)
                          : wfk::launch_conv<128, true, false>(plan->params, plan->grid, s);
  } else {
    e = (plan->bn == 256) ? wfk::launch_conv<256, false, false>(plan->params, plan->grid, s)
                          : wfk::launch_conv<128, false, false>(plan->params, plan->grid, s);
  }
  wfk::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return wfk::fail(WFK_ERR_CUDA, "launch of conv_gemm_kernel failed: %s", cudaGetErrorString(e));
  return WFK_OK;
}

extern "C" void wfk_conv_plan_destroy(wfk_conv_plan* plan) { delete plan; }
