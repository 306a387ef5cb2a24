// cluster_occ_probe.cu -- how many thread-block clusters of 2 / 4 / 8 CTAs (1 CTA per SM: ~200 KB of dynamic shared memory)
// can be co-resident on this GPU?  A 4-CTA cluster would let two CTA pairs share one multicast weight tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probes/cluster_occ_probe tools/probes/cluster_occ_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(sms / cs * cs); lc.blockDim = dim3(320); lc.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    lc.attrs = &at; lc.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &lc);
    printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
