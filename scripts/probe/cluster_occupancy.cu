// Probe: how many thread-block clusters of size C (one CTA per SM: 200 KB of dynamic shared memory) are co-resident on this
// device, and which SMs they land on.   nvcc -arch=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__global__ void probe(int* smid, int spin) {
  extern __shared__ char buf[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (threadIdx.x == 12345) buf[0] = 1;
}

int main() {
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int* d;
  cudaMalloc(&d, 4096 * sizeof(int));
  for (int c : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16}) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(c * 64);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
    printf("cluster %2d: max active clusters %3d -> %3d SMs (%s)\n", c, n, n * c, cudaGetErrorString(e));
    if (e != cudaSuccess || n < 1) { cudaGetLastError(); continue; }
    cfg.gridDim = dim3(c * n);
    cudaMemset(d, 0xff, 4096 * sizeof(int));
    e = cudaLaunchKernelEx(&cfg, probe, d, 2000000);
    cudaDeviceSynchronize();
    std::vector<int> h(c * n);
    cudaMemcpy(h.data(), d, c * n * sizeof(int), cudaMemcpyDeviceToHost);
    std::vector<int> used(h);
    std::sort(used.begin(), used.end());
    used.erase(std::unique(used.begin(), used.end()), used.end());
    printf("   distinct SMs used: %zu;", used.size());
    std::vector<char> mark(256, 0);
    for (int v : used) if (v >= 0 && v < 256) mark[v] = 1;
    printf(" idle smids:");
    for (int i = 0; i < 148; ++i) if (!mark[i]) printf(" %d", i);
    printf("\n");
  }
  return 0;
}
