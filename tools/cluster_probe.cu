// How many thread-block clusters of size 1/2/4/8/16 can be co-resident on this GPU for a kernel shaped like tp_gemm2_kernel
// (384 threads, ~200 KB dynamic shared memory, 1 CTA/SM)?  Planning data for cross-pair TMA multicast (cluster 4 / 8).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/cluster_probe tools/cluster_probe.cu && build/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 1) shaped_like_gemm(int* out) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0 && out != nullptr) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  const int smem = 201 * 1024;
  cudaFuncSetAttribute(shaped_like_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(shaped_like_gemm, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 200);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, shaped_like_gemm, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d SMs busy (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
