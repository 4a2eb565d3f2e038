// Panel rendering for log_wandb_images (reference pipeline/helpers.py:155-225): the numeric part of the figure,
// i.e. everything before matplotlib lays out axes --
//   target_np = (target.clamp(0,1) * 255).astype(uint8)                      (helpers.py:181)
//   pred_np   = (predicted.clamp(0,1) * 255).astype(uint8)                   (helpers.py:182)
//   diff_np   = |target_np - pred_np|                                        (helpers.py:183)
//   imshow(target_np / pred_np, cmap = vil_cmap(), norm = BoundaryNorm)      (helpers.py:187-203; sevir.py:1237-1268)
//   imshow(diff_np, cmap = 'Reds', vmin = 0, vmax = 255)                     (helpers.py:207)
// Both colour maps act on a uint8 value, so each is a 256-entry RGBA table (built on the host exactly as matplotlib
// builds its lookup table: weatherforecastingtoolkit_b200/render.py). HBM-bound: 8 B read + 15 B written per pixel,
// one pass, 4 pixels per thread (float4 loads, 32-bit / 128-bit stores).
#include "internal.h"

namespace wfk {

__device__ __forceinline__ uint32_t quant_u8(float v) {
  // torch.clamp propagates NaN; numpy's NaN -> uint8 cast is undefined (0 on x86-64): pin it to 0.
  v = fminf(fmaxf(v, 0.f), 1.f);
  return (v == v) ? static_cast<uint32_t>(__fmul_rn(v, 255.f)) : 0u;  // truncation, like ndarray.astype
}

__global__ void __launch_bounds__(256) render_panels_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                            int64_t n4, int64_t n, const uint32_t* __restrict__ lut_vil,
                                                            const uint32_t* __restrict__ lut_diff,
                                                            uint8_t* __restrict__ tgt_u8, uint8_t* __restrict__ pred_u8,
                                                            uint8_t* __restrict__ diff_u8, uint32_t* __restrict__ tgt_rgba,
                                                            uint32_t* __restrict__ pred_rgba,
                                                            uint32_t* __restrict__ diff_rgba) {
  __shared__ uint32_t s_vil[256], s_diff[256];
  s_vil[threadIdx.x] = lut_vil[threadIdx.x];
  s_diff[threadIdx.x] = lut_diff[threadIdx.x];
  __syncthreads();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t base = i << 2;
    float p[4], t[4];
    if (base + 3 < n) {
      const float4 pv = __ldg(reinterpret_cast<const float4*>(pred) + i);
      const float4 tv = __ldg(reinterpret_cast<const float4*>(tgt) + i);
      p[0] = pv.x, p[1] = pv.y, p[2] = pv.z, p[3] = pv.w;
      t[0] = tv.x, t[1] = tv.y, t[2] = tv.z, t[3] = tv.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[j] = (base + j < n) ? pred[base + j] : 0.f;
        t[j] = (base + j < n) ? tgt[base + j] : 0.f;
      }
    }
    uint32_t pq[4], tq[4], dq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pq[j] = quant_u8(p[j]);
      tq[j] = quant_u8(t[j]);
      dq[j] = pq[j] > tq[j] ? pq[j] - tq[j] : tq[j] - pq[j];
    }
    if (base + 3 < n) {
      if (tgt_u8) reinterpret_cast<uint32_t*>(tgt_u8)[i] = tq[0] | (tq[1] << 8) | (tq[2] << 16) | (tq[3] << 24);
      if (pred_u8) reinterpret_cast<uint32_t*>(pred_u8)[i] = pq[0] | (pq[1] << 8) | (pq[2] << 16) | (pq[3] << 24);
      if (diff_u8) reinterpret_cast<uint32_t*>(diff_u8)[i] = dq[0] | (dq[1] << 8) | (dq[2] << 16) | (dq[3] << 24);
      if (tgt_rgba) reinterpret_cast<uint4*>(tgt_rgba)[i] = make_uint4(s_vil[tq[0]], s_vil[tq[1]], s_vil[tq[2]], s_vil[tq[3]]);
      if (pred_rgba) reinterpret_cast<uint4*>(pred_rgba)[i] = make_uint4(s_vil[pq[0]], s_vil[pq[1]], s_vil[pq[2]], s_vil[pq[3]]);
      if (diff_rgba) reinterpret_cast<uint4*>(diff_rgba)[i] = make_uint4(s_diff[dq[0]], s_diff[dq[1]], s_diff[dq[2]], s_diff[dq[3]]);
    } else {
      for (int j = 0; j < 4 && base + j < n; ++j) {
        if (tgt_u8) tgt_u8[base + j] = static_cast<uint8_t>(tq[j]);
        if (pred_u8) pred_u8[base + j] = static_cast<uint8_t>(pq[j]);
        if (diff_u8) diff_u8[base + j] = static_cast<uint8_t>(dq[j]);
        if (tgt_rgba) tgt_rgba[base + j] = s_vil[tq[j]];
        if (pred_rgba) pred_rgba[base + j] = s_vil[pq[j]];
        if (diff_rgba) diff_rgba[base + j] = s_diff[dq[j]];
      }
    }
  }
}

}  // namespace wfk

extern "C" int wfk_render_panels(const float* pred, const float* tgt, int64_t count, const uint8_t* lut_vil_rgba,
                                 const uint8_t* lut_diff_rgba, uint8_t* tgt_u8, uint8_t* pred_u8, uint8_t* diff_u8,
                                 uint8_t* tgt_rgba, uint8_t* pred_rgba, uint8_t* diff_rgba, void* stream) {
  WFK_ENTER(stream, pred);
  WFK_REQUIRE(pred && tgt && lut_vil_rgba && lut_diff_rgba, "null pointer");
  WFK_REQUIRE(count > 0, "empty problem");
  auto aligned = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  WFK_REQUIRE(aligned(pred, 16) && aligned(tgt, 16) && aligned(lut_vil_rgba, 4) && aligned(lut_diff_rgba, 4),
              "inputs must be 16-byte aligned");
  WFK_REQUIRE(aligned(tgt_u8, 4) && aligned(pred_u8, 4) && aligned(diff_u8, 4) && aligned(tgt_rgba, 16) &&
                  aligned(pred_rgba, 16) && aligned(diff_rgba, 16),
              "outputs must be 4-byte (u8) / 16-byte (rgba) aligned");
  const int64_t n4 = (count + 3) >> 2;
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(wfk::num_sms()) * 16;  // grid-stride: 8 CTAs x 2 waves per SM
  if (blocks > cap) blocks = cap;
  wfk::render_panels_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, tgt, n4, count, reinterpret_cast<const uint32_t*>(lut_vil_rgba), reinterpret_cast<const uint32_t*>(lut_diff_rgba),
      tgt_u8, pred_u8, diff_u8, reinterpret_cast<uint32_t*>(tgt_rgba), reinterpret_cast<uint32_t*>(pred_rgba),
      reinterpret_cast<uint32_t*>(diff_rgba));
  return wfk::launched("render_panels_kernel");
}
