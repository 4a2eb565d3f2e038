// Probe: sustained issue rate of FFMA (3 register operands), FFMA with an immediate / constant operand, and the packed
// FFMA2 (fma.rn.f32x2) on one B200 SM sub-partition set -- the numbers the metrics kernel's fp32-pipe bound is derived from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/fp32_rate scripts/probes/fp32_rate.cu
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kIters = 4096;
constexpr int kChains = 16;

template <int MODE>
__global__ void rate_kernel(float* out, float a, float b, long long* cycles) {
  float x[kChains];
  float2 x2[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) {
    x[i] = threadIdx.x * 1e-3f + i;
    x2[i] = make_float2(x[i], x[i] + 0.5f);
  }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);                       // 3 register operands
      if (MODE == 1) x[i] = fmaf(x[i], 0.99993f, 1.0e-4f);          // immediates
      if (MODE == 2) x2[i] = __ffma2_rn(x2[i], a2, b2);             // packed
      if (MODE == 3) x[i] = x[i] * a;                               // FMUL
      if (MODE == 4) x[i] = fmaxf(x[i], a);                         // FMNMX (alu pipe)
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i] + x2[i].x + x2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  rate_kernel<MODE><<<148, threads>>>(out, 0.99993f, 1.0e-4f, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<MODE><<<148, threads>>>(out, 0.99993f, 1.0e-4f, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double instr_per_warp = double(kIters) * kChains;
  const double warps = threads / 32.0;
  const double per_clk_sm = instr_per_warp * warps / double(h[0]);   // warp-instructions per clock per SM
  printf("%-22s threads/SM %4d: %.2f warp-instr/clk/SM (%.1f lanes/clk/SM%s), %.1f us\n", name, threads, per_clk_sm,
         per_clk_sm * 32, MODE == 2 ? ", x2 FMAs" : "", ms * 1e3);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("FFMA reg,reg,reg", threads);
    run<1>("FFMA imm", threads);
    run<2>("FFMA2 (f32x2)", threads);
    run<3>("FMUL", threads);
    run<4>("FMNMX", threads);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
