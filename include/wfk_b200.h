/*
 * wfk_b200.h -- C ABI of libwfk_b200.so: the B200 (sm_100a) kernels behind the Path-B latent
 * nowcast rollout and its scoring.
 *
 * The reference (Autobot37/weatherforecastingtoolkit) has NO native layer: its boundary for this
 * path is plain Python. Every entry point below therefore cites the
 * reference PyTorch call site it replaces (paths relative to the reference repo root). The
 * reference-side binding a maintainer would add is the ctypes shim shown in INTEGRATION.md
 * (shipped as weatherforecastingtoolkit_b200/_cabi.py).
 *
 * Conventions: plain pointers and sizes only; all data pointers are DEVICE pointers to contiguous
 * buffers owned by the caller unless a parameter says "host"; `stream` is a cudaStream_t passed as
 * void*; every call only enqueues work on that stream (no hidden synchronisation) and returns 0 on
 * success or a negative wfk_status. There is no CPU fallback: an unsupported shape is an error.
 */
#ifndef WFK_B200_H
#define WFK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFK_ABI_VERSION 3

enum wfk_status {
  WFK_OK = 0,
  WFK_ERR_INVALID = -1,     /* bad argument / unsupported shape                      */
  WFK_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed (see wfk_last_error) */
  WFK_ERR_NO_DEVICE = -3,   /* no sm_100 device visible                              */
  WFK_ERR_NOT_INIT = -4,    /* wfk_init has not been called for the targeted device  */
  WFK_ERR_NONFINITE = -5    /* an activation left the fp16 range (inf / NaN detected) */
};

const char* wfk_strerror(int status);
const char* wfk_last_error(void);     /* detail string of the most recent failure (thread-local) */
int wfk_abi_version(void);
/* Register `device` with the library: check compute capability 10.x, create its context, resolve
 * cuTensorMapEncodeTiled. May be called for any number of devices, from any thread; the caller's current
 * device is left unchanged. Every other entry point takes its device from its `stream` argument (plan creation:
 * from the operand pointers), makes it current for the duration of the call and restores the caller's. */
int wfk_init(int device);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t wfk_launch_count(void);
/* Non-finite guard. The kernels that read GroupNorm statistics (wfk_gn_table, wfk_groupnorm_apply) or write model
 * outputs (the encoder's moments, the decoded frames) set a per-device host-mapped flag when they meet inf / NaN --
 * which is what an fp16 activation beyond 65504 becomes one layer later. Returns WFK_OK or WFK_ERR_NONFINITE (detail in
 * wfk_last_error); makes no CUDA call, so it is definitive only for work the caller has synchronised. reset != 0
 * clears the flag. The reference (fp32 / TF32) has no such failure mode: resnet.py / vae.py run in the fp32 range. */
int wfk_nonfinite_status(int device, int reset);

/* ---------------------------------------------------------------------------------------------
 * a1  VIL frame staging.  Replaces SEVIRDataLoader.preprocess_data_dict + change_layout
 *     (pipeline/datasets/sevir/sevir.py:626-666, 88-101; uint8->float cast at :587-592):
 *     out[n,t,0,h,w] = fl32(1/255) * (float)in[n,h,w,t], NHWT uint8 -> N T C H W.
 *     out_dtype: 0 = float32, 1 = float16 (round-to-nearest-even of the fp32 value).
 */
int wfk_stage_vil_u8(const uint8_t* nhwt, int n, int h, int w, int t, void* out_ntchw, int out_dtype,
                     void* stream);

/* 8(f).2  Event windowing + staging in one pass.  Replaces the sequence slicing of SEVIRDataLoader._idx_sample /
 *     _sequent_sample (pipeline/datasets/sevir/sevir.py:851-889: event_batch[e, :, :, s*stride : s*stride+seq_len])
 *     followed by preprocess_data_dict + change_layout.  events: [num_events, h, w, t_raw] uint8 resident on the device
 *     (SEVIR VIL events: t_raw = 49); windows: DEVICE int32 [n][2] = (event index, first raw frame); output
 *     [n, t, 1, h, w] as wfk_stage_vil_u8.  Window bounds (event < num_events, t0 + t <= t_raw) are the caller's
 *     contract (checked by the Python wrapper). */
int wfk_stage_vil_windows(const uint8_t* events, int num_events, int h, int w, int t_raw, const int32_t* windows, int n,
                          int t, void* out_ntchw, int out_dtype, void* stream);

/*     Same with the rescale constants of preprocess_data_dict spelled out: out = fl32(scale) * ((float)u8 + fl32(offset))
 *     -- rescale '01' = (1/255, 0), rescale 'sevir' = (1/47.54, -33.44) (sevir.py:44-63, 640-662). */
int wfk_stage_vil_windows_ex(const uint8_t* events, int num_events, int h, int w, int t_raw, const int32_t* windows, int n,
                             int t, float scale, float offset, void* out_ntchw, int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  Latent predictor.  Replaces the residual framing + nn.Linear(13*C, 12*C) + permutes of
 *     Model.validation_step (experiments/v1_experiments/pretrained_ae_linear_sevir/train.py:
 *     67, 101-113).  lat [B, t_in+t_out, C, HW] fp32 (HW = latent pixels); weight [t_out*C, t_in*C],
 *     bias [t_out*C] fp32.  Writes pred [B, t_out, C, HW] (= Linear(inp - last) + last) and, when
 *     non-NULL, tgt [B, t_out, C, HW] (copy of lat[:, t_in:]) and loss_sums[2] (double:
 *     sum((pred-tgt)^2), element count) -- the val_loss of train.py:109 is loss_sums[0]/loss_sums[1].
 *     loss_sums must be zeroed by the caller.
 */
int wfk_predict_linear(const float* lat, const float* weight, const float* bias, int b, int t_in, int t_out,
                       int c, int hw, float* pred, float* tgt, double* loss_sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 8(f).1  DLinear latent predictor.  Replaces moving_avg / series_decomp / DLinear.forward and the
 *     residual framing around it (experiments/v1_experiments/pretrained_ae_dlinear_sevir/train.py:
 *     21-99, 179-192; individual=True: ../pretrained_ae_dlinear_ind/train.py, experiments/ae_s2/
 *     train.py:55-133; the (t,c)-interleaved 52->48 form: ../pretrained_ae_dlinear_indc_indp/train.py:
 *     56-99, 185-186).
 *     framed != 0: x = latents [nb, (seq_len+pred_len)/group frames, ...] viewed as [nb][seq_len+pred_len]
 *       [channels] (x_batch_stride elements between sequences); the last input frame is subtracted
 *       before and added back after the predictor; writes pred [nb, pred_len, channels], optional tgt
 *       (same shape) and loss_sums[2] (double: sum((pred-tgt)^2) in residual space, element count;
 *       caller-zeroed).  group = 1: one series per (latent channel, pixel); group = C: series over
 *       the interleaved (t, c) axis, one per pixel.
 *     framed == 0: plain DLinear.forward, x [nb, seq_len, channels] -> pred [nb, pred_len, channels].
 *     individual == 0: w_* [pred_len, seq_len], b_* [pred_len]; else w_* [channels, pred_len, seq_len],
 *     b_* [channels, pred_len] (the stacked nn.ModuleList).  kernel_size odd (AvgPool1d window).
 */
int wfk_dlinear(const float* x, int64_t x_batch_stride, const float* w_seasonal, const float* b_seasonal,
                const float* w_trend, const float* b_trend, int nb, int seq_len, int pred_len, int channels,
                int group, int kernel_size, int individual, int framed, float* pred, float* tgt,
                double* loss_sums, void* stream);

/* 8(f).4  ConvModel latent compressor, one launch for the whole network.  Replaces ConvEncoder / ConvDecoder / ConvModel
 *     .forward (experiments/v1_experiments/pretrained_ae_convae_sevir/train.py:58-143) and the HuberLoss of its
 *     validation_step (train.py:160, 193-194).  x [n, cin, 48, 48] fp32 latent frames; weights: HOST array of 34 DEVICE
 *     fp32 pointers in module order: conv0 (w, b); LayerNorm (w, b) of conv0, down1..3, up1..3; down1..3 (w, b);
 *     to_latent (w, b); to_reconstruction (w, b); up1..3 (w, b); conv_out (w, b).  Writes z [n, latent_dim],
 *     recon [n, cin, 48, 48] and, when non-NULL, huber_sums[2] (double: sum of Huber(recon, x), element count;
 *     caller-zeroed). */
int wfk_convmodel_forward(const float* x, int n, int cin, int latent_dim, const float* const* weights, float* z,
                          float* recon, double* huber_sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a10-a14  Fused skill-score pass.  Replaces the 41 passes of calc_metrics
 *     (pipeline/metrics.py:86-133): clamp(0,1) (:92-93); _hit_miss_fa_cn counts (:9-16) at the
 *     fp32 thresholds for pools none / avg4 / avg16 (csi :43-54, hss :56-69, avg_pool2d :46-50);
 *     crps == MAE at one member (:18-41); torchmetrics SSIM (:71-75) and per-frame PSNR (:77-84).
 */
#define WFK_MAX_THRESHOLDS 8
#define WFK_NUM_POOLS 3 /* pool 1, 4, 16 */

typedef struct wfk_metric_partials {
  /* exact integer contingency counts over all frames: [pool][threshold][tp, fn, fp, tn] */
  int64_t counts[WFK_NUM_POOLS][WFK_MAX_THRESHOLDS][4];
  int64_t n_elems[WFK_NUM_POOLS]; /* pixels (pooled cells) counted per pool                  */
  int64_t n_frames;               /* frames scored                                            */
  double abs_sum[WFK_NUM_POOLS];  /* sum |pred - tgt| of the (pooled) clamped fields         */
  double sq_sum;                  /* sum (pred - tgt)^2, pool 1                               */
  double ssim_sum;                /* sum over frames of the per-frame mean SSIM               */
  double psnr_sum;                /* sum over frames of 10*log10(range^2 / mse_frame)         */
  double reserved[2];
} wfk_metric_partials;

size_t wfk_metrics_workspace_bytes(int frames, int h, int w);
/* pred, tgt: [frames, h, w] fp32. clamp01 != 0 applies calc_metrics' clamp(0,1) (metrics.py:92-93)
 * on load; 0 scores the values as given (the stand-alone csi/hss/ssim/psnr/crps functions).
 * thresholds: HOST array of fp32 values (already rounded the way torch rounds the Python-float
 * threshold, SURVEY hazard H2). out: DEVICE struct, fully overwritten. workspace: DEVICE scratch
 * of wfk_metrics_workspace_bytes. */
int wfk_metrics(const float* pred, const float* tgt, int frames, int h, int w, const float* thresholds,
                int n_thresholds, int clamp01, wfk_metric_partials* out, void* workspace,
                size_t workspace_bytes, void* stream);

/* The rest of pipeline/metrics.py's public surface (not used by calc_metrics' fixed sweep): any pooling window and
 * pool_type='max' in csi / hss / crps (metrics.py:22-32, 43-50, 56-63), ensemble forecasts (pred.ndim == 6: Gaussian CRPS
 * over n > 1 members :33-41, pred.mean(dim=1) in calc_metrics :94).  pool_kind: 0 none (scale 1), 1 avg (F.avg_pool2d:
 * sequential row-major fp32 window sum, one division), 2 max (F.max_pool2d); stride == scale; cells = floor(h/scale) x
 * floor(w/scale).
 *   wfk_pooled_counts: counts [n_thresholds][4] int64 (tp, fn, fp, tn; exact), sums[0] += sum |p - t| over the pooled
 *     cells, sums[1] += number of cells.  thresholds: HOST fp32 array.  counts / sums: DEVICE, caller-zeroed.
 *   wfk_crps_ensemble: pred [b, n, tc, h, w], tgt [b, tc, h, w]; sums[0] += sum of the per-cell CRPS terms, sums[1] +=
 *     cells (the reference returns their ratio).
 *   wfk_ensemble_mean: out [b, inner] = mean over the n members of pred [b, n, inner] (fp32, members summed in order). */
int wfk_pooled_counts(const float* pred, const float* tgt, int frames, int h, int w, int pool_kind, int scale,
                      const float* thresholds, int n_thresholds, int clamp01, int64_t* counts, double* sums, void* stream);
int wfk_crps_ensemble(const float* pred, const float* tgt, int b, int n, int tc, int h, int w, int pool_kind, int scale,
                      int clamp01, double* sums, void* stream);
int wfk_ensemble_mean(const float* pred, int b, int n, int64_t inner, int clamp01, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3-a6, a9  Autoencoder building blocks (activations NHWC, 16 bits per value on device: fp16 by default; the
 *     entry points with a `bf16` argument / wfk_conv_desc.operand_bf16 read and write bf16 instead -- the whole chain
 *     of one model must use one format).
 *
 * Implicit-GEMM convolution / GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces nn.Conv2d 3x3 / 1x1 (resnet.py:405,421,452; vae.py:24,68,103,148), Downsample2D
 * (resnet.py:181-190), Upsample2D (resnet.py:108-143), the residual add (resnet.py:493) and the
 * attention linears / baddbmm / bmm (attention.py:146-176), and accumulates the GroupNorm
 * statistics of its own output (resnet.py:403,419) in the epilogue.
 *
 * One launch computes, for every output tile (128 pixels x BN channels),
 *     D = sum over taps  A_src[pixel + (dx,dy), k-range] * B_src[slab, n, k-range]
 * followed by  out = D + bias (+ residual), optional fp16 / fp32 stores and per-(frame, group)
 * sum / sum-of-squares accumulation.
 */
typedef struct wfk_tap {
  int8_t dx, dy;    /* pixel offset added to the tile origin (x, y coordinate of the A view)   */
  int8_t q;         /* coordinate of the auxiliary A dimension (0 unless a phase/parity view)   */
  int8_t src;       /* which A / B source pair (0 or 1)                                         */
  int16_t c_off;    /* starting element along the innermost A dimension                         */
  int16_t b_slab;   /* slab (3rd coordinate) of the B source                                    */
  int16_t kblocks;  /* number of 64-element K blocks this tap contributes                       */
  int16_t reserved;
} wfk_tap;

/* A view: 5-D, innermost first: (k, x, q, y, frame); strides in BYTES (stride[0] is implied). */
typedef struct wfk_view5 {
  const void* ptr;
  int64_t dim[5];
  int64_t stride[5];
} wfk_view5;
/* B view: 3-D: (k, n, slab). */
typedef struct wfk_view3 {
  const void* ptr;
  int64_t dim[3];
  int64_t stride[3];
} wfk_view3;

#define WFK_MAX_TAPS 40

/* Epilogue activations (conv-GEMM `act` / `act2`, the direct edge kernels' `act`). */
enum wfk_act {
  WFK_ACT_NONE = 0,
  WFK_ACT_LEAKY_RELU = 1, /* slope = act_slope (nn.LeakyReLU(0.2), losses/model.py:121)          */
  WFK_ACT_GELU = 2,       /* exact erf form (nn.GELU(), ae_64x8x8_lin.py:15, ae_vit.py:112)       */
  WFK_ACT_SIGMOID = 3,    /* nn.Sigmoid (ae_64x8x8_lin.py:85)                                      */
  WFK_ACT_SILU = 4,
  /* Row-wise softmax split over three GEMM launches (attention.py:163-176: baddbmm -> softmax(fp32) -> bmm) so that
   * the fp32 score matrix never exists in memory. Plain one-tap GEMM descriptors only (`act` field, fp16 operands):
   *   ROW_MAX : no tensor output; row_out[row][slot] = max over the slot's columns of D          (pass 1: Q K^T)
   *   ROW_EXP : m = max_slot row_in[row][slot]; out_h = exp2((D - m) * row_scale);
   *             row_out[row][slot] = sum over the slot's columns of that value (fp32)             (pass 2: Q K^T again)
   *   ROW_NORM: out = D / (sum_slot row_in[row][slot]) + shift2[c]                                (pass 3: P V)
   * row = frame * out_rows * out_cols + pixel; every (row, slot) is written by exactly one thread, in a fixed
   * order, so the results are deterministic. row_ld = number of slots = 2 * ceil(n_total / 256) of the GEMM that WRITES
   * them (128 instead of 256 when n_total < 256 and not a multiple of 256); ROW_NORM reads its producer's row_ld. */
  WFK_ACT_ROW_MAX = 5,
  WFK_ACT_ROW_EXP = 6,
  WFK_ACT_ROW_NORM = 7
};

typedef struct wfk_conv_desc {
  wfk_view5 a[2];
  wfk_view3 b[2];
  int32_t n_frames;      /* tiles are enumerated per frame                                      */
  int32_t tile_h, tile_w;/* extent of the tile coordinate space (rows, cols of output pixels per phase) */
  int32_t n_total;       /* GEMM N (output channels); multiple of 8 (of the N tile when stats)   */
  int32_t num_phases;    /* 1, or 4 for the sub-pixel (nearest x2 upsample) decomposition        */
  int32_t taps_per_phase;
  wfk_tap taps[WFK_MAX_TAPS]; /* [phase][tap]                                                    */
  int32_t a_frame_mul;   /* A frame coordinate = frame * a_frame_mul  (0: A shared by all frames) */
  int32_t b_frame_mul;   /* B slab coordinate  = tap.b_slab + frame * b_frame_mul                */
  const float* bias;     /* [n_total] or NULL                                                   */
  const void* residual;  /* fp16, same addressing as out, or NULL                               */
  void* out_h;           /* fp16 output or NULL                                                 */
  float* out_f;          /* fp32 output or NULL                                                 */
  double* stats;         /* [n_frames][n_total/cpg][2] accumulators (caller-zeroed) or NULL     */
  int32_t out_rows, out_cols; /* output tensor spatial dims                                     */
  int32_t out_sy, out_sx;/* output pixel = (y*out_sy + (phase>>1), x*out_sx + (phase&1))         */
  int32_t ldc;           /* output / residual channel pitch in elements                         */
  int32_t cpg;           /* channels per GroupNorm group (4, 8 or 16) when stats != NULL        */
  int32_t operand_bf16;  /* 0: fp16 operands and fp16 activation I/O (default), 1: bf16 for both      */
  const void* gn_table;  /* optional fused GroupNorm+SiLU on A source 0 (3x3 stride-1 convs only):
                            [n_frames][cin] float2 (scale, shift) from wfk_gn_table, or NULL          */
  /* Alternative to gn_table: derive (scale, shift) inside the kernel from the producer's raw statistics
   * (same arithmetic as wfk_gn_table; no separate launch). gn_stats: [n_frames][gn_groups][2] double (sum, sum of
   * squares over cin/gn_groups channels x tile_h*tile_w pixels); gamma, beta: [cin]. */
  const double* gn_stats;
  const float* gn_gamma;
  const float* gn_beta;
  float gn_eps;
  int32_t gn_groups;
  /* Extended epilogue (eval-mode BatchNorm folded into weights/bias by the caller):
   *   v = act(D + bias (+ residual))            -> out_h / out_f / stats
   *   out2_h = act2(scale2[c] * v + shift2[c])  -> a second fp16 tensor, same addressing as out_h
   * e.g. a pre-activation bottleneck's x (raw, the residual) and gelu(bn1(x)) (the next conv's input)
   * from one pass (ae_64x8x8_lin.py:14-23); LeakyReLU after conv+BN (losses/model.py:126-139). */
  int32_t act;           /* wfk_act on the primary result                                       */
  int32_t act2;          /* wfk_act on the secondary result                                     */
  float act_slope;       /* LeakyReLU slope                                                     */
  void* out2_h;          /* fp16 or NULL                                                        */
  const float* scale2;   /* [n_total] or NULL (= 1)                                             */
  const float* shift2;   /* [n_total] or NULL (= 0)                                             */
  /* WFK_ACT_ROW_* only (see wfk_act): per-row partials, [rows][row_ld] fp32 */
  const float* row_in;
  float* row_out;
  int32_t row_ld;
  float row_scale;       /* ROW_EXP: log2(e) * softmax scale                                    */
} wfk_conv_desc;

typedef struct wfk_conv_plan wfk_conv_plan;
int wfk_conv_plan_create(const wfk_conv_desc* desc, wfk_conv_plan** out);
int wfk_conv_plan_run(const wfk_conv_plan* plan, void* stream);
void wfk_conv_plan_destroy(wfk_conv_plan* plan);

/* GroupNorm apply (+ optional SiLU).  Replaces nn.GroupNorm(32, C, eps=1e-6) + SiLU
 * (resnet.py:457-458, 479-485; vae.py:82-83, 162-163; attention.py:141).  x, out: [n, hw, c] fp16;
 * stats: [n][groups][2] double (sum, sum of squares over the group's c/groups * hw values). */
int wfk_groupnorm_apply(const void* x, const double* stats, const float* gamma, const float* beta, int n,
                        int hw, int c, int groups, float eps, int apply_silu, void* out, int bf16, void* stream);

/* Per-(frame, channel) scale/shift of a GroupNorm whose apply (+SiLU) is fused into the consuming
 * convolution's operand staging: table[n][c] = (rstd*gamma, beta - mean*rstd*gamma). */
int wfk_gn_table(const double* stats, const float* gamma, const float* beta, int n, int hw, int c, int groups,
                 float eps, void* table, void* stream);

/* Direct 3x3 (pad 1, stride 1) convolution for tiny input-channel counts.  Replaces
 * encoder.conv_in (vae.py:24) and post_quant_conv + decoder.conv_in (autoencoder_kl.py:87,
 * vae.py:103).  in: [n, cin, h, w] fp32 NCHW; optional pre 1x1 conv (pre_w [cin,cin], pre_b [cin])
 * applied to in-bounds taps; w: [cout, cin, 3, 3] fp32; out: [n, h, w, cout] fp16 + stats. */
int wfk_conv3x3_small_cin(const float* in, int n, int cin, int h, int w, const float* pre_w,
                          const float* pre_b, const float* weight, const float* bias, int cout, void* out,
                          double* stats, int cpg, void* stream);

/* The same layers on the tensor cores (mma.sync m16n8k16, fp16 operands / fp32 accumulate): K = (cin + ones_plane)
 * * 9 <= 48. weight_h: [ceil(K/16)*16][cout] fp16, row k = ci*9 + tap, zero rows beyond K; with ones_plane != 0 an
 * extra constant-one input plane (1 inside the image, 0 in the padding) carries the bias of a 1x1 convolution folded
 * in front (post_quant_conv, autoencoder_kl.py:87). cout a multiple of 128; cpg 4, 8 or 16 when stats != NULL. */
int wfk_conv3x3_stem_tc(const float* in, int n, int cin, int h, int w, int ones_plane, const void* weight_h,
                        const float* bias, int cout, void* out, double* stats, int cpg, int bf16, void* stream);

/* Direct 3x3 (pad 1, stride 1) convolution for tiny output-channel counts, optional post 1x1.
 * Replaces decoder.conv_out (vae.py:148) and encoder.conv_out + quant_conv (vae.py:68,
 * autoencoder_kl.py:82).  in: [n, h, w, cin] fp16 (already normalised + SiLU); weight_h:
 * [cout][9][cin] fp16; bias [cout]; post_w [cout,cout], post_b [cout] or NULL; out [n, cout, h, w] fp32. */
int wfk_conv3x3_small_cout(const void* in, int n, int h, int w, int cin, const void* weight_h,
                           const float* bias, int cout, const float* post_w, const float* post_b, float* out,
                           void* stream);

/* Same with an output activation (WFK_ACT_NONE or WFK_ACT_SIGMOID): PosAwareAE_TF's
 * dec[-1] Conv2d(128, 1, 3, padding=1) + Sigmoid (pipeline/models/ae_64x8x8_lin.py:83-85, 105). */
int wfk_conv3x3_small_cout_act(const void* in, int n, int h, int w, int cin, const void* weight_h,
                               const float* bias, int cout, const float* post_w, const float* post_b, int act,
                               float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a16 / a17  Stem and head kernels of PosAwareAE_TF and NLayerDiscriminator (eval-mode BatchNorm
 * folded into weight / bias by the caller).
 *
 * Conv2d(1, cout, 4, stride 2, padding 1) (+BN) + activation on an fp32 frame: enc[0].down of
 * PosAwareAE_TF (ae_64x8x8_lin.py:31-34) and main[0..1] of NLayerDiscriminator (losses/model.py:121).
 * in [n, 1, h, w] fp32 (h, w even); weight [16][cout] fp32 (tap r*4+s major); bias [cout];
 * out [n, h/2, w/2, cout] fp16 NHWC = act(conv + bias); optional out2 = act2(scale2*out + shift2). */
int wfk_conv4x4s2_c1in(const float* in, int n, int h, int w, const float* weight, const float* bias, int cout,
                       int act, float act_slope, void* out, int act2, const float* scale2, const float* shift2,
                       void* out2, void* stream);

/* Conv2d(cin, 1, kernel 1, padding pad): NLayerDiscriminator's logit head (losses/model.py:149-150; the
 * reference pads a 1x1 convolution, so the border ring of the output is the bare bias).
 * in [n, h, w, cin] fp16 NHWC; weight [cin] fp32; out [n, 1, h+2*pad, w+2*pad] fp32. */
int wfk_conv1x1_cout1(const void* in, int n, int h, int w, int cin, const float* weight, float bias, int pad,
                      float* out, void* stream);

/* Reductions behind the discriminator losses (losses/contperceptual.py:19-23 hinge_d_loss; g_loss =
 * -mean(logits_fake), experiments/v1_experiments/ae_gan_kl/train.py:84-85): sums[0] += sum x,
 * sums[1] += sum relu(1 - x), sums[2] += sum relu(1 + x) over count values (double, caller-zeroed). */
int wfk_logit_sums(const float* x, int64_t count, double* sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a18  Token-path kernels of AE_ViT_2048 (pipeline/models/ae_vit.py:84-162); the dense linears run on
 * wfk_conv_plan_* GEMMs.
 *
 * Conv2d(1, D, 16, 16) / ConvTranspose2d(D, 1, 16, 16) as GEMMs over 16x16 patches (ae_vit.py:100, 135,
 * 141-143, 158-159): img [n, 1, h, w] fp32 <-> rows [n*(h/16)*(w/16)][256] (fp16 for patchify, fp32 for
 * unpatchify), element (r, s) of a patch at column r*16 + s. */
int wfk_patchify16(const float* img, int n, int h, int w, void* rows_h, void* stream);
int wfk_unpatchify16(const float* rows_f, int n, int h, int w, float* img, void* stream);
/* nn.MultiheadAttention core (ae_vit.py:106-111, 127-132): qkv [n*tokens][3*d_model] fp16 (q | k | v, heads are
 * contiguous 64-wide slices) -> out [n*tokens][d_model] fp16 = softmax(q k^T / 8) v per (image, head).
 * tokens <= 64, d_model = heads * 64. */
int wfk_mha_small(const void* qkv, int n, int tokens, int d_model, int heads, void* out, void* stream);
/* GlobalCrossEncode core (ae_vit.py:23-41): q_scaled [heads][d_latent/heads] fp32 = q_proj(query_vec) * scale
 * (input independent); kv [n*tokens][2*d_latent] fp16 (k | v) -> out [n][d_latent] fp16. */
int wfk_cross_encode_attn(const float* q_scaled, const void* kv, int n, int tokens, int d_latent, int heads,
                          void* out, void* stream);
/* nn.LayerNorm(d, eps) over rows of an fp32 matrix (the post-norm residual sums, ae_vit.py:106-111):
 * out_h fp16 and / or out_f fp32. */
int wfk_layernorm_rows(const float* x, int64_t rows, int d, const float* gamma, const float* beta, float eps,
                       void* out_h, float* out_f, void* stream);
/* out[b, l, :] = vec[b, :] + pos[l, :] (fp32 in, fp16 out): GlobalCrossDecode with a single key/value token
 * (ae_vit.py:44-82: softmax over one key is exactly 1) followed by + pos_embed (ae_vit.py:152). */
int wfk_bcast_add_rows(const float* vec, const float* pos, int n, int tokens, int d, void* out, void* stream);

/* Decoder tail, fused: GroupNorm(groups, eps) + SiLU + conv3x3(c -> 1, pad 1).  Replaces
 * conv_norm_out + conv_act + conv_out of Decoder.forward (vae.py:162-164).  x: [n, h, w, c] fp16 raw
 * stream; stats as wfk_groupnorm_apply; weight: [9][c] fp32 (tap-major); out: [n, 1, h, w] fp32. */
int wfk_gn_silu_conv3x3_c1(const void* x, const double* stats, const float* gamma, const float* beta, int n,
                           int h, int w, int c, int groups, float eps, const float* weight, float bias,
                           float* out, int bf16, void* stream);

/* Row softmax: probs[r, :] = softmax(scale * scores[r, :]) in fp32, stored fp16.
 * Replaces torch.softmax(attention_scores.float(), dim=-1) (attention.py:171); `scale` is the
 * baddbmm alpha (attention.py:148, 168). */
int wfk_softmax_rows(const float* scores, int64_t rows, int cols, float scale, void* probs, int bf16, void* stream);

/* f.4  ConvAttnModel latent compressor (experiments/v1_experiments/pretrained_ae_convattn_ae_sevir/train.py:58-170),
 *      whole network in one launch, one CTA per latent frame: x [n, cin, 48, 48] fp32 -> z [n, latent_dim] and
 *      recon [n, cin, 48, 48]; huber_sums (optional, double[2]) += (sum of HuberLoss(recon, x) terms, element count)
 *      as in the experiment's validation_step (train.py:206-209).  mode 0: encode + decode (forward), 1: encode only
 *      (z written), 2: decode only (z read; x may be NULL).  Fixed by the kernel: 48 x 48 input, embed 128, 8 heads,
 *      feed-forward 512; runtime: cin <= 8, layers <= 8, latent_dim <= 512.  weights: 29 + 30 * layers DEVICE fp32
 *      pointers, 16-byte aligned, in the order listed in predictors.py (ConvAttnModel._weight_pointers); the matrices
 *      of the token GEMMs (self-attention in/out projections, linear1, linear2, the pooling key/value projection) are
 *      passed TRANSPOSED ([in, out]), everything else in the PyTorch layout. */
int wfk_convattn_forward(const float* x, int n, int cin, int layers, int latent_dim, const float* const* weights,
                         int num_weights, float* z, float* recon, double* huber_sums, int mode, void* stream);

/* f.3  Validation panels.  Replaces the numpy + matplotlib colour mapping of log_wandb_images
 *      (pipeline/helpers.py:155-225) with vil_cmap() (pipeline/datasets/sevir/sevir.py:1237-1268):
 *        tgt_u8 = (uint8)(clamp(tgt,0,1) * 255), pred_u8 likewise (truncation), diff_u8 = |tgt_u8 - pred_u8|,
 *        tgt_rgba = lut_vil[tgt_u8], pred_rgba = lut_vil[pred_u8], diff_rgba = lut_diff[diff_u8].
 *      pred, tgt: DEVICE fp32 [count]; lut_*: DEVICE uint8 [256][4]; outputs DEVICE uint8 [count] / [count][4], each
 *      may be NULL.  One pass, HBM-bound (8 B read + 15 B written per pixel). */
int wfk_render_panels(const float* pred, const float* tgt, int64_t count, const uint8_t* lut_vil_rgba,
                      const uint8_t* lut_diff_rgba, uint8_t* tgt_u8, uint8_t* pred_u8, uint8_t* diff_u8,
                      uint8_t* tgt_rgba, uint8_t* pred_rgba, uint8_t* diff_rgba, void* stream);

/* a7  DiagonalGaussianDistribution arithmetic (pipeline/models/autoencoderkl/distributions.py:26-42) in one pass:
 * moments [n, 2*lc, hw] fp32 -> logvar = clamp(moments[:, lc:], -30, 20), std = exp(0.5*logvar), var = exp(logvar)
 * (each [n, lc, hw], any may be NULL) and, when noise != NULL, sample = mean + std * noise. */
int wfk_gaussian_posterior(const float* moments, int n, int lc, int hw, float* logvar, float* std, float* var,
                           const float* noise, float* sample, void* stream);

/* [n, hw, c] fp32 -> [n, c, hw] fp32: the conv-GEMM's NHWC result to the model-facing NCHW moments tensor
 * (AutoencoderKL.encode returns [B, 2*latent_channels, h, w], autoencoder_kl.py:80-84). */
int wfk_nhwc_to_nchw_f32(const float* in, int n, int hw, int c, float* out, void* stream);

/* fp32 -> fp16 conversion of weights at pack time (device to device). */
int wfk_f32_to_f16(const float* in, int64_t n, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WFK_B200_H */
