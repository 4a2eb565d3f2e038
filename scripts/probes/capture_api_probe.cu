// Probe: which runtime calls are legal while a stream is being captured into a CUDA graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probes/capture_api_probe scripts/probes/capture_api_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int* p) { if (p) *p = 1; }
#define TRY(name, expr)                                                        \
  do {                                                                         \
    cudaError_t e = (expr);                                                    \
    cudaStreamCaptureStatus st;                                                \
    cudaStreamIsCapturing(s, &st);                                             \
    printf("%-28s -> %s, capture status %d\n", name, cudaGetErrorName(e), (int)st); \
  } while (0)
int main() {
  cudaStream_t s;
  cudaStreamCreate(&s);
  int* d;
  cudaMalloc(&d, 4);
  for (int variant = 0; variant < 5; ++variant) {
    cudaGraph_t g = nullptr;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    int dev = -1;
    cudaPointerAttributes attr;
    if (variant == 0) TRY("cudaStreamGetDevice", cudaStreamGetDevice(s, &dev));
    if (variant == 1) TRY("cudaGetDevice", cudaGetDevice(&dev));
    if (variant == 2) TRY("cudaPointerGetAttributes", cudaPointerGetAttributes(&attr, d));
    if (variant == 3) TRY("cudaGetLastError", cudaGetLastError());
    if (variant == 4) TRY("cudaFuncSetAttribute", cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024));
    k<<<1, 1, 0, s>>>(d);
    TRY("  launch", cudaGetLastError());
    cudaError_t e = cudaStreamEndCapture(s, &g);
    printf("  end capture: %s graph %p\n", cudaGetErrorName(e), (void*)g);
    cudaGetLastError();
  }
  return 0;
}
