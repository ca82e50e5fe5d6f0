// Probe: how many clusters of size 2/4/8 (1 CTA per SM, ~210 KB smem) can be resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) { extern __shared__ char s[]; if (out) out[0] = s[0]; }
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  size_t smem = 210 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim = {(unsigned)cs, 1, 1};
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
