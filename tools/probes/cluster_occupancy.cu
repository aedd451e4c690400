// How many clusters of 2 / 4 / 8 CTAs (1 CTA per SM, ~210 KB dynamic smem, 320 threads) can be co-resident?
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 210 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 148);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = cs, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cfg.attrs = &at, cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
