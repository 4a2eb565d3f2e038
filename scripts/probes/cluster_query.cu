// How many 2- / 4- / 8-CTA clusters of a 1-CTA-per-SM kernel (220 KB dynamic shared memory, 480 threads) can be
// co-resident on this GPU? (Sizing question for a TMA-multicast weight ring across CTA pairs.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148 / cs * cs);
    cfg.blockDim = dim3(480);
    cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
