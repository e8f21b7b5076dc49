// How many clusters of 2 / 4 / 8 CTAs (1 CTA per SM: 200 KB of dynamic shared memory, 384 threads) can be resident on the
// device at once?  Decides whether an 8-CTA cluster GEMM (4 CTA pairs sharing a multicast W tile) could use all SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_occ tools/probe/cluster_occ.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 60);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d = %3d of %d SMs (%s)\n", cs, n, n * cs, sms, cudaGetErrorString(e));
  }
  return 0;
}
