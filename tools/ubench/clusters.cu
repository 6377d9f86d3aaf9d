// How many thread-block clusters of each size can be co-resident on this GPU (1 CTA/SM at large smem)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out) { extern __shared__ char s[]; if (threadIdx.x == 0 && out) out[blockIdx.x] = s[0]; }
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : {32 * 1024, 100 * 1024, 204 * 1024, 227 * 1024}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t q{}; q.gridDim = dim3(cs * 64); q.blockDim = dim3(288); q.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      q.attrs = a; q.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &q);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
  }
  return 0;
}
