// How many clusters of 1/2/4/8 CTAs (576 threads, ~196 KB shared memory each: the sweep kernels' footprint)
// can be co-resident on this GPU?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k(int *p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main()
{
    const int smem = 176 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int cs = 1; cs <= 16; cs *= 2) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1; cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        cudaGetLastError();
    }
    return 0;
}
