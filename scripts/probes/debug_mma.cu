// Hardware probe (test-only entry point): does tcgen05.mma accept a K-major SWIZZLE_128B A operand whose
// start address is an arbitrary multiple of 128 B inside a TMA-written (k, P, ROWS) halo tile, with the
// 8-row groups `P` rows apart (SBO = P*128 B)? This is what a 3x3 convolution needs to reuse ONE staged
// halo tile for all nine taps. Result decides the design of the halo conv kernel.
#include "internal.h"
#include "ptx.cuh"

namespace wfk {

struct DebugMmaParams {
  CUtensorMap a_map;  // 3-D (k=64, P, ROWS) fp16
  CUtensorMap b_map;  // 2-D as 3-D (k=64, 128, 1)
  int pitch, rows, r, s, base_offset;
  float* d;  // [128][128]
};

__global__ void __launch_bounds__(128, 1) debug_shifted_mma_kernel(const __grid_constant__ DebugMmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = p.pitch * p.rows * 128;
  uint8_t* sb = smem + ((a_bytes + 1023) & ~1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 128 * 128);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, a_bytes + 128 * 128);
    tma_load_3d(smem, &p.a_map, bar, 0, 0, 0);
    tma_load_3d(sb, &p.b_map, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(smem) + (p.r * p.pitch + p.s) * 128;
    const uint32_t b_addr = smem_u32(sb);
    const uint32_t idesc = umma_idesc_f16(128, 128, 0);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = 0;
      ad |= static_cast<uint64_t>(((a_addr + k * 32) & 0x3FFFF) >> 4);
      ad |= static_cast<uint64_t>(1) << 16;
      ad |= static_cast<uint64_t>((p.pitch * 128) >> 4) << 32;  // SBO = halo row pitch
      ad |= static_cast<uint64_t>(1) << 46;
      ad |= static_cast<uint64_t>(p.base_offset & 7) << 49;
      ad |= static_cast<uint64_t>(2) << 61;
      umma_f16(tmem_base, ad, umma_desc_sw128(b_addr + k * 32), idesc, k > 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t rr[32];
  for (int c0 = 0; c0 < 128; c0 += 32) {
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, rr);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(rr[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

}  // namespace wfk

extern "C" int wfk_debug_shifted_mma(const void* a_halo, int pitch, int rows, const void* b, int r, int s,
                                     int base_offset, float* d, void* stream) {
  WFK_ENTER_STREAM(stream);
  WFK_REQUIRE(a_halo && b && d, "null pointer");
  WFK_REQUIRE(pitch >= 10 && pitch <= 32 && rows >= 18 && rows <= 64, "bad halo shape");
  wfk::DebugMmaParams p{};
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(pitch), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[2] = {128, static_cast<cuuint64_t>(pitch) * 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(pitch), static_cast<cuuint32_t>(rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = wfk::g_encode_tiled(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(a_halo), dims,
                                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return wfk::fail(WFK_ERR_CUDA, "encode A failed: %d", (int)rc);
  cuuint64_t bdims[3] = {64, 128, 1};
  cuuint64_t bstr[2] = {128, 128 * 128};
  cuuint32_t bbox[3] = {64, 128, 1};
  rc = wfk::g_encode_tiled(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(b), bdims, bstr, bbox, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return wfk::fail(WFK_ERR_CUDA, "encode B failed: %d", (int)rc);
  p.pitch = pitch;
  p.rows = rows;
  p.r = r;
  p.s = s;
  p.base_offset = base_offset;
  p.d = d;
  const size_t smem = 1024 + ((static_cast<size_t>(pitch) * rows * 128 + 1023) & ~size_t(1023)) + 128 * 128 + 64;
  static wfk::PerDeviceOnce attr_once;
  if (wfk::PerDeviceOnce::Lock attr_lock{attr_once}; attr_lock.needed()) {
    WFK_CUDA_CHECK(cudaFuncSetAttribute(wfk::debug_shifted_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_lock.finished();
  }
  wfk::debug_shifted_mma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return wfk::launched("debug_shifted_mma_kernel");
}
