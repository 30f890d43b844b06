// Probe: how many thread-block clusters of size 8 / 9 / 16 with ~190 KB of shared memory per CTA can be resident on
// this GPU at once (cudaOccupancyMaxActiveClusters), i.e. whether 16 clusters of 9 CTAs (one 48x48 patch each) fit.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void dummy(int* out) {
  extern __shared__ unsigned char smem[];
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) smem[0] = 1;
  cl.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<int>(cl.num_blocks());
}

int main() {
  int dev = 0;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, dev);
  printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
  const int smem = 190 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {2, 4, 8, 9, 12, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 16);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %2d: max active clusters = %d (%s)\n", cs, n, cudaGetErrorString(e));
    int* d; cudaMalloc(&d, 4);
    e = cudaLaunchKernelEx(&cfg, dummy, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    printf("            launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d);
  }
  return 0;
}
