// Validates, on one CTA, every tcgen05 assumption the reduced-precision strict kernel relies on:
//   * shared-memory operand layout (K-major, no swizzle: 8x16-byte core matrices, SBO = 128 B,
//     LBO = rows/8 * 128 B) and the 64-bit matrix descriptor built for it;
//   * the kind::f16 instruction descriptor (FP16 operands, FP32 accumulate, M = 128, N = 160);
//   * accumulator placement in tensor memory (row m -> lane m, column n -> column n) as seen by
//     tcgen05.ld.32x32b from the warp that owns lane quadrant warp % 4;
//   * tcgen05.st / tcgen05.ld round trip, column-unaligned loads;
//   * cycle counts: 30 MMAs (3 passes x 10 K-steps) + commit, and a full 128 x 160 FP32 read-back.
// Prints max |D - A B^T| against the host and the timings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, NN = 160, K = 160;
constexpr int SBO = 128;
constexpr int LBO_A = (M / 8) * 128;    // 2048
constexpr int LBO_B = (NN / 8) * 128;   // 2560
constexpr int A_BYTES = (K / 8) * LBO_A;   // 40960
constexpr int B_BYTES = (K / 8) * LBO_B;   // 51200

__host__ __device__ inline int off_kmajor(int row, int k, int lbo) {
    return (k >> 3) * lbo + (row >> 3) * SBO + (row & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // layout type 0: no swizzle
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__global__ void __launch_bounds__(160, 1) k_probe(const __half* gA, const __half* gB, float* gD, float* gRT,
                                                  float* gUn, long long* gClk) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + A_BYTES;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operands -> shared memory in the K-major core-matrix layout (generic-proxy stores)
    for (int i = tid; i < M * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sA + off_kmajor(r, k, LBO_A)) = gA[i];
    }
    for (int i = tid; i < NN * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sB + off_kmajor(r, k, LBO_B)) = gB[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // instruction descriptor: D = F32 (1<<4), A = B = F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 4 && lane == 0) {
        t0 = clock64();
        for (int pass = 0; pass < 3; ++pass)
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t ad = make_desc(smem_u32(sA) + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bd = make_desc(smem_u32(sB) + ks * 2 * LBO_B, LBO_B, SBO);
                umma_f16(tmem, ad, bd, idesc, (pass | ks) ? 1u : 0u);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
        t1 = clock64();
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(ok)
                : "r"(smem_u32(&bar)), "r"(0)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 && lane == 0) t2 = clock64();

    if (warp < 4) {
        const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
        const int row = 32 * warp + lane;
        long long r0 = clock64();
        for (int c = 0; c < NN; c += 8) {
            uint32_t v[8];
            tmem_ld8(lane_base + c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 8; ++e) gD[row * NN + c + e] = __uint_as_float(v[e]);
        }
        long long r1 = clock64();
        if (tid == 0) gClk[3] = r1 - r0;
        // store / load round trip into columns 256..263, then a column-unaligned load at 257
        uint32_t w[8];
        for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(1000.f * row + e);
        tmem_st8(lane_base + 256, w);
        for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(1000.f * row + 8 + e);
        tmem_st8(lane_base + 264, w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t v[8];
        tmem_ld8(lane_base + 256, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) gRT[row * 8 + e] = __uint_as_float(v[e]);
        tmem_ld8(lane_base + 257, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) gUn[row * 8 + e] = __uint_as_float(v[e]);
        // read throughput: 20 x8 loads back to back, one wait
        long long q0 = clock64();
        float s = 0.f;
        for (int rep = 0; rep < 4; ++rep) {
            uint32_t a[8], b[8], c2[8], d[8];
            tmem_ld8(lane_base + 0 + rep * 32, a);
            tmem_ld8(lane_base + 8 + rep * 32, b);
            tmem_ld8(lane_base + 16 + rep * 32, c2);
            tmem_ld8(lane_base + 24 + rep * 32, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 8; ++e) s += __uint_as_float(a[e]) + __uint_as_float(b[e]) + __uint_as_float(c2[e]) + __uint_as_float(d[e]);
        }
        long long q1 = clock64();
        if (s == 1.2345f) gD[0] = s;
        if (tid == 0) gClk[4] = q1 - q0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        if (lane == 0) { gClk[0] = t1 - t0; gClk[1] = t2 - t0; gClk[2] = 0; }
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}

// TMEM read / write throughput with `blockDim.x / 32` warps (4 per lane quadrant share the lanes).
// mode 0: x8 loads, 5 in flight; mode 1: x32 loads; mode 2: x8 stores.
__global__ void __launch_bounds__(1024, 1) k_tmem_bw(int mode, int reps, long long* gClk, float* sink) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16);
    const int part = warp >> 2, parts = blockDim.x >> 7;
    const int cols = 480 / parts;                 // columns this warp covers per rep (multiple of 40)
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int c = 0; c < cols; c += 40) {
            const uint32_t base = lane_base + part * cols + c;
            if (mode == 0) {
                uint32_t a[8], b[8], c2[8], d[8], e2[8];
                tmem_ld8(base, a); tmem_ld8(base + 8, b); tmem_ld8(base + 16, c2); tmem_ld8(base + 24, d); tmem_ld8(base + 32, e2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int e = 0; e < 8; ++e)
                    acc += __uint_as_float(a[e]) + __uint_as_float(b[e]) + __uint_as_float(c2[e]) + __uint_as_float(d[e]) + __uint_as_float(e2[e]);
            } else if (mode == 1) {
                uint32_t a[32];
                tmem_ld32(base, a);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int e = 0; e < 32; ++e) acc += __uint_as_float(a[e]);
            } else {
                uint32_t w[8];
                for (int e = 0; e < 8; ++e) w[e] = __float_as_uint(acc + e);
                tmem_st8(base, w); tmem_st8(base + 8, w); tmem_st8(base + 16, w); tmem_st8(base + 24, w); tmem_st8(base + 32, w);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (acc == 1.2345f) sink[0] = acc;
    if (tid == 0) gClk[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
}

int main() {
    std::vector<__half> hA(M * K), hB(NN * K);
    std::vector<float> fA(M * K), fB(NN * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < NN * K; ++i) { float v = (rand() % 2001 - 1000) / 3000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD, *dRT, *dUn; long long* dClk;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dD, M * NN * 4); cudaMalloc(&dRT, M * 8 * 4); cudaMalloc(&dUn, M * 8 * 4); cudaMalloc(&dClk, 8 * 8);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * NN * 4); cudaMemset(dClk, 0, 64);
    const int smem = A_BYTES + B_BYTES;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 160, smem>>>(dA, dB, dD, dRT, dUn, dClk);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
    std::vector<float> hD(M * NN), hRT(M * 8), hUn(M * 8);
    long long clk[8];
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hRT.data(), dRT, hRT.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hUn.data(), dUn, hUn.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(clk, dClk, 64, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < NN; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
            ref *= 3.0;   // three accumulating passes over the same operands
            maxerr = fmax(maxerr, fabs(ref - hD[m * NN + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    int rt_bad = 0, un_bad = 0;
    for (int m = 0; m < M; ++m)
        for (int e = 0; e < 8; ++e) {
            if (hRT[m * 8 + e] != 1000.f * m + e) ++rt_bad;
            if (hUn[m * 8 + e] != 1000.f * m + e + 1) ++un_bad;
        }
    printf("{\"umma_max_abs_err\": %.3e, \"umma_max_ref\": %.3e, \"tmem_roundtrip_bad\": %d, \"unaligned_ld_bad\": %d, "
           "\"clk_issue_30mma\": %lld, \"clk_30mma_done\": %lld, \"clk_readback_160cols_waited\": %lld, "
           "\"clk_read_128cols_pipelined\": %lld}\n",
           maxerr, maxref, rt_bad, un_bad, clk[0], clk[1], clk[3], clk[4]);
    // tensor-memory bandwidth: bytes per clock per SM for 4, 8, 16 warps
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8, 16}) {
            const int reps = 50;
            cudaMemset(dClk, 0, 64);
            k_tmem_bw<<<1, warps * 32>>>(mode, reps, dClk, dD);
            err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
            cudaMemcpy(clk, dClk, 8, cudaMemcpyDeviceToHost);
            const int parts = warps / 4, cols = (480 / parts) / 40 * 40;
            const double bytes = (double)reps * cols * parts * 128 * 4 * (mode == 1 ? 0.8 : 1.0);
            printf("{\"tmem_bw_mode\": \"%s\", \"warps\": %d, \"bytes_per_clk\": %.1f}\n",
                   mode == 0 ? "ld.x8" : mode == 1 ? "ld.x32(40-col stride)" : "st.x8", warps, bytes / (double)clk[0]);
        }
    return 0;
}
