// Device-side building blocks shared by the FP64 tensor-pipe kernels (strict Wieber rollout, Herdt tick):
// DMMA.8x8x4 wrapper, bulk async copy + mbarrier plumbing, named group barriers, the double-buffered
// contraction loop.  Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace lipm {
namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Bulk asynchronous copy global -> shared (TMA unit, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                          unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :
        : "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

constexpr int kGroupCols = 8;                    // one DMMA column block

// Named barriers: id 0 is __syncthreads; group g uses id 1 + g.  GT = threads of a solver group.
template <int GT>
__device__ __forceinline__ void group_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(GT) : "memory");
}
template <int GT>
__device__ __forceinline__ bool group_sync_and(int id, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %2, 0;\n"
        "bar.red.and.pred q, %1, %3, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(out)
        : "r"(id), "r"((uint32_t)pred), "n"(GT)
        : "memory");
    return out != 0;
}

// Non-negative doubles order like their high words; 20 mantissa bits are plenty for a stop test,
// and the maxima then run on the integer pipe instead of competing with DMMA for the FP64 pipe.
__device__ __forceinline__ uint32_t mag_hi(double x) { return (uint32_t)__double2hiint(x) & 0x7fffffffu; }

// Box projection by plain compare-select (fmin/fmax would add NaN-quieting code to every call).
__device__ __forceinline__ double clipd(double x, double lo, double hi) {
    const double r = x < lo ? lo : x;
    return r > hi ? hi : r;
}

// acc[r] += M[row block r of this warp] * R[:, this group's 8 columns]   (NR row blocks, all k).
// Fragments are double buffered: the loads of k-block kb+1 are in flight while kb is multiplied.
// (No __restrict__ on bBase: R is rewritten between calls and the loads must stay behind the
// group barrier that publishes it.)
// KBC > 0: compile-time trip count, fully unrolled (straight-line code keeps ptxas' warp
// convergence check -- YIELD / WARPSYNC.ALL -- out of the inner loop); KBC == 0: run-time count.
// STREAM: the matrix is not resident in shared memory (horizons beyond the resident limit); its
// fragments are read straight from global memory (they stay in L2: one coalesced 256-byte line
// per warp and fragment, read-only path).
template <int NR, int NRB, int KBC, bool STREAM>
__device__ __forceinline__ void contract(double (&acc)[NRB][2], const double* aBase,
                                         const int aStride, const volatile double* bBase, const int KB_rt) {
    const int KB = KBC > 0 ? KBC : KB_rt;
    double a0[NR], a1[NR], b0, b1;
    auto lda = [&](int off) -> double { return STREAM ? __ldg(aBase + off) : aBase[off]; };
    b0 = bBase[0];
#pragma unroll
    for (int r = 0; r < NR; ++r) a0[r] = lda(r * aStride);
    // With a single row block the two DMMAs of a trip would form one dependent chain on the same accumulator;
    // odd K blocks then go to a second accumulator (summed at the end).
    double odd0 = 0.0, odd1 = 0.0;
    auto trip = [&](int kb) {
        b1 = bBase[(kb + 1) * 4];
#pragma unroll
        for (int r = 0; r < NR; ++r) a1[r] = lda(r * aStride + (kb + 1) * 32);
#pragma unroll
        for (int r = 0; r < NR; ++r) dmma884(acc[r][0], acc[r][1], a0[r], b0);
        if (kb + 2 < KB) {
            b0 = bBase[(kb + 2) * 4];
#pragma unroll
            for (int r = 0; r < NR; ++r) a0[r] = lda(r * aStride + (kb + 2) * 32);
        }
        if (NR == 1) {
            dmma884(odd0, odd1, a1[0], b1);
        } else {
#pragma unroll
            for (int r = 0; r < NR; ++r) dmma884(acc[r][0], acc[r][1], a1[r], b1);
        }
    };
    if (KBC > 0) {
#pragma unroll
        for (int kb = 0; kb < KBC; kb += 2) trip(kb);
    } else {
#pragma unroll 1
        for (int kb = 0; kb < KB; kb += 2) trip(kb);     // KB = NP/4 is even: NP is a multiple of 8
    }
    if (NR == 1) { acc[0][0] += odd0; acc[0][1] += odd1; }
}
template <int NR, int NRB, int KBC, bool STREAM>
struct ContractDispatch {
    static __device__ __forceinline__ void run(int nrb, double (&acc)[NRB][2], const double* aBase, int aStride,
                                               const volatile double* bBase, int KB) {
        if (nrb == NR) contract<NR, NRB, KBC, STREAM>(acc, aBase, aStride, bBase, KB);
        else ContractDispatch<NR - 1, NRB, KBC, STREAM>::run(nrb, acc, aBase, aStride, bBase, KB);
    }
};
template <int NRB, int KBC, bool STREAM>
struct ContractDispatch<0, NRB, KBC, STREAM> {
    static __device__ __forceinline__ void run(int, double (&)[NRB][2], const double*, int, const volatile double*, int) {}
};

}  // namespace
}  // namespace lipm
