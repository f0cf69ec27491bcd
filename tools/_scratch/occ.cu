#include <cstdio>
#include <cuda_runtime.h>
template <int R> __global__ void __launch_bounds__(256, 2) k(float* p) {
    extern __shared__ float sm[];
    float acc[R];
    for (int i = 0; i < R; ++i) acc[i] = p[i * 256 + threadIdx.x];
    for (int j = 0; j < 100; ++j) for (int i = 0; i < R; ++i) acc[i] = acc[i] * acc[(i + 1) % R] + sm[(i * 7 + j) % 64];
    float s = 0; for (int i = 0; i < R; ++i) s += acc[i];
    p[threadIdx.x] = s;
}
template <int R> void test() {
    auto kern = k<R>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    for (int kb : {48, 64, 90, 96, 100, 110, 112, 113}) {
        size_t smem = (size_t)kb * 1024;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int blocks = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 256, smem);
        printf("regs %d smem %d KB -> blocks/SM %d\n", fa.numRegs, kb, blocks);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("regsPerSM %d regsPerBlock %d smemPerSM %zu smemPerBlockOptin %zu reserved %zu\n", p.regsPerMultiprocessor, p.regsPerBlock, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.reservedSharedMemPerBlock);
    test<40>(); test<100>();
    return 0;
}
