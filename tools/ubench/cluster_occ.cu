// Probe: how many thread-block clusters of size 8 / 16 are co-resident on this GPU for a kernel using S KB of dynamic shared memory?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  for (int cl : {2, 4, 8, 16}) for (int kb : {64, 128, 200, 215}) {
    size_t smem = (size_t)kb * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cl * 32); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d smem %3d KB: max active clusters %d (%s) -> %d SMs\n", cl, kb, n, cudaGetErrorString(e), n * cl);
  }
  return 0;
}
