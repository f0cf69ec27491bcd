// Register-resident FP64 throughput probes for the roofline denominator (B200, sm_100a):
//   dmma : independent DMMA.8x8x4 accumulator chains per warp (the tensor path the strict kernel uses)
//   dfma : independent DFMA chains per thread (the vector FP64 pipe)
// Prints TFLOP/s for several warps-per-SM settings.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, double seed) {
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, double seed) {
    double c[CHAINS];
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    double a = 1.0 + seed * 1e-12, b = seed * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    printf("{\"device\": \"%s\", \"sms\": %d, \"probes\": [\n", prop.name, sms);
    const int warps_list[] = {4, 8, 12, 16, 32};
    bool first = true;
    for (int w : warps_list) {
        float ms = time_ms([&] { k_dmma<8><<<sms, 32 * w>>>(out, iters, 1.0); });
        double flops = 2.0 * 256 * 8.0 * iters * w * sms;
        printf("%s {\"kind\": \"dmma884\", \"warps_per_sm\": %d, \"chains\": 8, \"tflops\": %.3f}", first ? "" : ",\n", w, flops / ms / 1e9);
        first = false;
        ms = time_ms([&] { k_dfma<8><<<sms, 32 * w>>>(out, iters * 8, 1.0); });
        flops = 2.0 * 32 * 8.0 * iters * 8 * w * sms;
        printf(",\n {\"kind\": \"dfma\", \"warps_per_sm\": %d, \"chains\": 8, \"tflops\": %.3f}", w, flops / ms / 1e9);
    }
    // dependent-issue latency: one warp per SM sub-partition with 1..10 independent chains, and
    // 2 or 3 warps per sub-partition with 5 chains each (the strict kernel's situation)
#define CHAIN_PROBE(CH, W)                                                                          \
    {                                                                                               \
        float ms = time_ms([&] { k_dmma<CH><<<sms, 32 * (W)>>>(out, iters, 1.0); });              \
        double flops = 2.0 * 256 * (double)(CH) * iters * (W) * sms;                                \
        printf(",\n {\"kind\": \"dmma884\", \"warps_per_sm\": %d, \"chains\": %d, \"tflops\": %.3f}", (W), (CH), \
               flops / ms / 1e9);                                                                   \
    }
    CHAIN_PROBE(1, 4) CHAIN_PROBE(2, 4) CHAIN_PROBE(3, 4) CHAIN_PROBE(4, 4) CHAIN_PROBE(5, 4) CHAIN_PROBE(6, 4)
    CHAIN_PROBE(10, 4) CHAIN_PROBE(5, 8) CHAIN_PROBE(5, 12) CHAIN_PROBE(10, 8)
    printf("\n]}\n");
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
