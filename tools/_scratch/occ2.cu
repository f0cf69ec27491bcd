#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int F> __global__ void __launch_bounds__(256, 2) k(float* p) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint32_t* base = reinterpret_cast<uint32_t*>(sm);
    float acc[90];
    for (int i = 0; i < 90; ++i) acc[i] = p[i * 256 + threadIdx.x];
    if (F & 1) __syncthreads();
    if (F & 2) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base[0]) : "memory");
    }
    if (F & 4) {
        uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); acc[0] += r;
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (F & 8) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(base[0]) : "memory");
    }
    for (int j = 0; j < 100; ++j) for (int i = 0; i < 90; ++i) acc[i] = acc[i] * acc[(i + 1) % 90] + (float)sm[(i * 7 + j) % 64];
    float s = 0; for (int i = 0; i < 90; ++i) s += acc[i];
    p[threadIdx.x] = s;
}
template <int F> void test() {
    auto kern = k<F>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    size_t smem = 97720;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int blocks = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 256, smem);
    printf("features %d regs %d -> blocks/SM %d\n", F, fa.numRegs, blocks);
}
int main() { test<0>(); test<1>(); test<2>(); test<3>(); test<4>(); test<5>(); test<8>(); test<9>(); test<13>(); return 0; }
