// Strict (box-constrained) receding-horizon rollout, reduced-precision path on CTA PAIRS: the ADMM contraction as
// tcgen05.mma.cta_group::2 (two SMs of a TPC work on one 128-column tile), FP16 hi/lo operands, FP32 accumulators
// in tensor memory.
//
// Same algorithm, outputs, precision scheme and reference code as lipm_rollout_strict_f16.cu
// (zmp_controller.py:173-195 strict QP, :199 state update, :81-106 windowing / padding / push; v = M r as three
// kind::f16 passes M_hi r_hi + M_hi r_lo + M_lo r_hi; FP32 ADMM state in tick-relative coordinates; FP64 walk state,
// applied sample clamped to the FP64 box).  What changes is how a 128-column unit is mapped to the machine:
//
//   * a unit is solved by a CLUSTER of two CTAs.  D[column, row] = A[column, k] * B[row, k] with M = 128 columns,
//     N = 160 rows, K = 160 is ONE pair instruction: each CTA supplies the right-hand sides of ITS 64 columns (A,
//     20 KB per FP16 image) and HALF of the shared matrix (B rows 80 r .. 80 r + 79 of [M | M G Px], 25 KB per
//     image), the tensor cores of both SMs exchange the B halves, and each CTA receives D for its own 64 columns.
//     30 pair MMAs (M = 128, N = 160, K = 16) cost 40 cycles each: the same tensor throughput per SM as the
//     single-CTA tile, with HALF the shared memory (131 KB -> 91 KB per CTA) and half the tensor memory;
//   * the accumulator of a CTA uses the "2 x 2" tensor-memory layout: lanes 0..63 hold rows 0..79 of the CTA's 64
//     columns, lanes 64..127 hold rows 80..159, in 80 tensor-memory columns.  D, t and lo take 240 of 512 columns;
//   * that leaves room for TWO CTAs per SM (8 warps each): while one waits for its contraction, the other runs its
//     FP32 epilogue -- the overlap the single-CTA kernel could not have (its second column set needed 640 tensor-
//     memory columns and 262 KB of shared memory) -- and for small batches a 20-warp CTA per SM instead, so that
//     a 4096-scenario sweep (64 units) keeps 128 SMs busy instead of 64 and each column's 160 rows are spread over
//     10 threads instead of 4;
//   * per iteration the two CTAs meet once: each publishes its right-hand sides, tells the other "published, and
//     all my columns are done / not done" (remote flag + remote mbarrier arrive with cluster-scope release), and
//     both wait for the other's word.  The leader then issues the 30 MMAs and commits them to an mbarrier in BOTH
//     CTAs (multicast).  Columns still freeze individually, so results do not depend on which other columns share
//     the pair, the CTA, the batch or the GPU.
#include "lipm_internal.cuh"
#include "lipm_tcgen05.cuh"

#include <cuda_fp16.h>

namespace lipm {

namespace {

constexpr int kPairCols = 64;                         // columns per CTA (half of the 128-column unit)
constexpr int kHalfRows = kF16Tile / 2;               // 80 accumulator rows per lane half
constexpr int kHalfChunks = kHalfRows / 8;            // 10 8-row chunks per half
constexpr int kLboAp = (kPairCols / 8) * 128;         // 1024: A image, 64 rows per K chunk
constexpr int kLboBp = (kHalfRows / 8) * 128;         // 1280: B half image, 80 rows per K chunk
constexpr int kImgA = (kF16Tile / 8) * kLboAp;        // 20480 B per FP16 image of A
constexpr int kImgB = (kF16Tile / 8) * kLboBp;        // 25600 B per FP16 image of the B half
constexpr uint32_t kColD = 0, kColT = kHalfRows, kColLo = 2 * kHalfRows;   // tensor-memory column offsets (256 allocated)
constexpr uint32_t kTmemCols = 256;

struct PairParams {
    ColumnBatch cb;
    int N;
    int max_iter;
    float rho, two_rho, alpha, qhalf;
    float eps_stop;          // eps / max(1 + alpha, rho alpha)
    double viol_tol;
    double T, c2, c3, hg, px02, pu00;
    const unsigned char* bimg;   // FP16 hi image then lo image of B (160 rows, plan layout)
    const double* gpx;           // [NP,3]
    unsigned int* tile_counter;
    unsigned long long* stats;
    unsigned int num_units;
};

template <int kParts>
struct PairControl {
    unsigned long long bar_load, bar_mma;
    uint32_t tmem_base;
    int unit, max_ticks;
    int other_ticks;             // written by the other CTA of the pair
    int other_unit;              // unit index handed over by the leader
    uint32_t other_word;         // written by the other CTA: (exchange number << 1) | "all my columns are done"
    float gpx[kF16Tile * 3];
    double X[kPairCols][3];
    uint32_t Res[2 * kParts][kPairCols];
    uint32_t Qm[2 * kParts][kPairCols];
};

constexpr size_t kSmemOperandsP = 2 * (size_t)kImgB + 2 * (size_t)kImgA;    // 92,160

using namespace tc;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_store(uint32_t remote_addr, int v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// One word per exchange instead of an mbarrier round trip: the sender stores (sequence << 1 | done) with cluster-scope
// release, the receiver polls its own copy with cluster-scope acquire.
__device__ __forceinline__ void remote_store_release(uint32_t remote_addr, uint32_t v) {
    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t load_acquire_cluster(const void* local) {
    uint32_t v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(local)) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Issued by a whole, converged warp: one elected lane sends the instruction, and because every operand is
// warp-uniform the descriptors stay in uniform registers (issuing from `if (lane == 0)` instead costs an
// ELECT + five R2UR.BROADCAST per MMA and made the contraction issue-bound: 72 cycles per 40-cycle MMA).
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, leader;\n\telect.sync _|leader, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@leader tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
    asm volatile(
        "{\n\t.reg .pred leader;\n\telect.sync _|leader, 0xffffffff;\n\t"
        "@leader tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(
            smem_u32(bar)),
        "h"((unsigned short)3)
        : "memory");
}

// Eight consecutive contraction slots of one column, split into FP16 hi + lo, as two 16-byte stores into the
// operand images (K chunk kc, column s of this CTA: byte offset kc * LBO + s * 16).
__device__ __forceinline__ void publish_rhs_pair(unsigned char* sAh, unsigned char* sAl, int kc, int s, const float (&r)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(r[2 * e], r[2 * e + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(r[2 * e] - back.x, r[2 * e + 1] - back.y);
        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const int off = kc * kLboAp + s * 16;
    *reinterpret_cast<uint4*>(sAh + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sAl + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// kParts = row parts per lane half: 5 (20 warps, 16 rows per thread, one CTA per SM: small batches) or
// 2 (8 warps, 40 rows per thread, two CTAs per SM: the epilogue of one overlaps the contraction of the other).
template <int kParts>
__global__ void __launch_bounds__(128 * kParts, kParts == 2 ? 2 : 1)
k_strict_rollout_f16p(const PairParams P) {
    constexpr int kThreads = 128 * kParts;
    constexpr int kChunks = kHalfChunks / kParts;        // 8-row chunks per thread
    static_assert(kChunks * kParts == kHalfChunks, "parts must divide the 10 chunks of a lane half");
    using Control = PairControl<kParts>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sBh = smem;
    unsigned char* sBl = smem + kImgB;
    unsigned char* sAh = smem + 2 * kImgB;
    unsigned char* sAl = sAh + kImgA;
    Control& S = *reinterpret_cast<Control*>(smem + kSmemOperandsP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, part = warp >> 2;
    const int tl = quad * 32 + lane;                 // tensor-memory lane
    const int s = tl & (kPairCols - 1);              // column of this CTA
    const int hf = tl >> 6;                          // lane half: rows 0..79 or 80..159
    const int chunk0 = part * kChunks;               // first 8-row chunk (within the half) of this thread
    const int row_base = hf * kHalfRows + chunk0 * 8;   // first accumulator row of this thread
    const bool state_chunk_owner = (hf == 1) && (part == kParts - 1);   // its LAST chunk is rows 152..159: the state slots
    const bool head = (hf == 0) && (part == 0);      // owns preview row 0 and the outputs of the column
    const int N = P.N;
    const ColumnBatch& B = P.cb;
    const uint32_t rank = cluster_rank();
    const uint32_t other = rank ^ 1u;

    // ---- one-time setup: barriers, tensor memory, this CTA's half of the shared operand B, G Px ----------------
    if (tid == 0) {
        mbar_init(&S.bar_load, 1);
        mbar_init(&S.bar_mma, 1);
        S.other_word = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < kF16Tile * 3; i += kThreads) S.gpx[i] = (i < N * 3) ? (float)P.gpx[i] : 0.f;
    for (int i = tid; i < 2 * kImgA / 16; i += kThreads) reinterpret_cast<uint4*>(sAh)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        // rows 80 rank .. 80 rank + 79 of the plan's 160-row image: 1280 of every 2560 bytes (one K chunk)
        mbar_expect_tx(&S.bar_load, 2u * kImgB);
        for (int img = 0; img < 2; ++img)
            for (int kc = 0; kc < kF16Tile / 8; ++kc)
                bulk_load((img ? sBl : sBh) + kc * kLboBp, P.bimg + (size_t)img * kF16OperandBytesB + (size_t)kc * (2 * kLboBp) + rank * kLboBp,
                          kLboBp, &S.bar_load);
    }
    mbar_wait(&S.bar_load, 0);
    cluster_sync_all();                               // both CTAs have their barriers and tensor memory before anyone signals
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, S.tmem_base, 0) + kColD;     // accumulator of the pair instruction
    const uint32_t tmem = S.tmem_base + ((uint32_t)(32 * quad) << 16);   // this warp's lane quadrant
    // D = FP32, A = B = FP16, both K-major, N = 160, M = 128 (the pair's tile)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kF16Tile >> 3) << 17) | ((uint32_t)(kF16Cols >> 4) << 24);
    const uint32_t other_word_addr = map_to_cta(smem_u32(&S.other_word), other);
    const uint32_t other_ticks_addr = map_to_cta(smem_u32(&S.other_ticks), other);
    const uint32_t other_unit_addr = map_to_cta(smem_u32(&S.other_unit), other);

    // tensor memory starts undefined: clear t and lo once so padding rows read as finite zeros
    {
        uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
            tmem_st8(tmem + kColT + (chunk0 + ch) * 8, z);
            tmem_st8(tmem + kColLo + (chunk0 + ch) * 8, z);
        }
        tmem_wait_st();
    }

    float hi[kChunks * 8];        // upper bounds of this thread's rows (tick-relative, FP32)
    uint32_t mma_parity = 0;
    unsigned int exchange = 0;    // exchanges with the other CTA so far (flag slot = exchange & 1)

    for (;;) {
        // ---- the pair fetches a 128-column unit (leader asks the queue, hands the index over) ---------------
        __syncthreads();
        if (tid == 0) {
            S.max_ticks = 0;
            if (rank == 0) {
                S.unit = (int)atomicAdd(P.tile_counter, 1u);
                remote_store(other_unit_addr, S.unit);
            }
        }
        cluster_sync_all();
        if (tid == 0 && rank != 0) S.unit = S.other_unit;
        __syncthreads();
        const unsigned int unit = (unsigned int)S.unit;
        if (unit >= P.num_units) break;
        const int c = (B.unit_order ? B.unit_order[unit] : (int)(P.num_units - 1u - unit)) * kF16Cols + (int)rank * kPairCols + s;

        // column description (every thread of a column decodes it; the head thread owns the outputs)
        const bool valid = c < B.ncols;
        const int axis = valid ? c / B.B : 0;
        int b = valid ? c - axis * B.B : 0;
        if (valid && B.order) b = B.order[b];
        int n = valid ? (B.n_steps ? B.n_steps[b] : B.n_uniform) : 0;
        n = n > B.n_max ? B.n_max : n;
        int nticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
        if (!valid || n < 1 || nticks < 0) nticks = 0;
        const int last = n - 1;
        const long long bz = B.bounds_index ? B.bounds_index[b] : ((B.bounds_batch == 1) ? 0 : b);
        const double* plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        const double* phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        double* ph = nullptr;
        double *pcom = nullptr, *pzmp = nullptr;
        uint8_t* pf = nullptr;
        double imp = 0.0, viol = 0.0, lean = 0.0;
        int ftick = -1, notconv = 0, iter_sum = 0;
        if (head) {
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (valid) {
                const double* init = (axis ? B.init[1] : B.init[0]) + (long long)b * 3;
                x0 = init[0]; x1 = init[1]; x2 = init[2];
                ph = axis ? B.hist[1] : B.hist[0];
                if (ph) ph += (long long)b * B.hist_scn_stride;
                pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
                pcom = B.com ? B.com + ((long long)b * B.n_max) * 2 + axis : nullptr;
                pzmp = B.zmp ? B.zmp + ((long long)b * B.n_max) * 2 + axis : nullptr;
                if (axis == 1 && B.impulse != nullptr) {
                    imp = B.impulse[b];
                    ftick = B.force_tick ? B.force_tick[b] : -1;
                }
                if (B.write_initial && nticks > 0) {
                    const double zmp = x0 - P.hg * x2;
                    const double below = plo[0] - zmp, above = zmp - phi[0];
                    viol = fmax(fmax(below, above), 0.0);
                    lean = fabs(x0 - zmp);
                    if (ph) { ph[0] = x0; ph[1] = x1; ph[2] = x2; }
                    if (pcom) pcom[0] = x0;
                    if (pzmp) pzmp[0] = zmp;
                    if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                              (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
                }
            }
            S.X[s][0] = x0; S.X[s][1] = x1; S.X[s][2] = x2;
            atomicMax(&S.max_ticks, nticks);
        }
        __syncthreads();
        if (tid == 0) remote_store(other_ticks_addr, S.max_ticks);
        cluster_sync_all();
        const int max_ticks = max(S.max_ticks, S.other_ticks);      // the pair steps its ticks together
        unsigned int unit_iters = 0;
        double origin = 0.0;          // origin of the tick-relative coordinates (FP64)
#ifdef LIPM_F16P_PROFILE
        long long pc_exch = 0, pc_mma = 0, pc_epi = 0, pc_pro = 0, pc_apply = 0, pc_t = 0;
#define PC_MARK(acc) do { const long long now__ = clock64(); acc += now__ - pc_t; pc_t = now__; } while (0)
        pc_t = clock64();
#else
#define PC_MARK(acc) do { } while (0)
#endif

        for (int i = 0; i < max_ticks; ++i) {
            const bool live = i < nticks;
            float thr = 0.f;
            {
                // ================= tick prologue =================
                const double x0d = S.X[s][0];
                const float x1f = (float)S.X[s][1], x2f = (float)S.X[s][2];
                const float shift = (float)(origin - x0d);     // re-reference the warm start
                origin = x0d;
                // phase 1: bound window of this tick (relative to the origin); warm start = previous tick's t moved
                // up one row, staged through the (idle) accumulator columns.  Row j + 1 of a thread's last row may
                // live in the other lane half (row 79 -> 80): that one value travels through shared memory (Qm is idle).
                if (i > 0 && hf == 1 && part == 0) {
                    uint32_t first[8];
                    tmem_ld8(tmem + kColT, first);             // rows 80..87 of the previous tick
                    tmem_wait_ld();
                    S.Qm[0][s] = first[0];
                }
                __syncthreads();
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    if (state_chunk_owner && ch == kChunks - 1) continue;      // slots 152..159: axis state
                    const int cb8 = (chunk0 + ch) * 8;                         // tensor-memory column of the chunk
                    const int jb = row_base + ch * 8;                          // accumulator row of the chunk
                    uint32_t lov[8], tv[8], tu[8];
                    if (i > 0) {
                        tmem_ld8(tmem + kColT + cb8 + 1, tv);      // rows jb+1 .. jb+8 of the previous tick
                        tmem_ld8(tmem + kColT + cb8, tu);          // (row N-1 keeps its own value)
                        tmem_wait_ld();
                        if (hf == 0 && cb8 + 8 == kHalfRows) tv[7] = S.Qm[0][s];   // row 80 lives in the other lane half
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = jb + e;
                        float l = 0.f, h = 0.f;
                        if (live && j < N) {
                            int smp = i + B.bound_index_bias + j;
                            smp = smp > last ? last : smp;
                            l = (float)(__ldg(plo + (long long)smp * B.bound_stride) - x0d);
                            h = (float)(__ldg(phi + (long long)smp * B.bound_stride) - x0d);
                        }
                        lov[e] = __float_as_uint(l);
                        hi[ch * 8 + e] = h;
                    }
                    tmem_st8(tmem + kColLo + cb8, lov);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = jb + e;
                        const float l = __uint_as_float(lov[e]), h = hi[ch * 8 + e];
                        float t;
                        if (!live || j >= N) t = 0.f;
                        else if (i == 0) t = 0.5f * (l + h);
                        else if (j + 1 < N) t = __uint_as_float(tv[e]) + shift;
                        else t = clipf(__uint_as_float(tu[e]) + shift, l, h);   // last row: keep the primal value, drop the dual
                        tv[e] = __float_as_uint(t);
                    }
                    tmem_st8(tmem + kColD + cb8, tv);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncthreads();          // every thread has read the old t before anyone overwrites it
                tc_fence_after();
                // phase 2: t <- staged value; |q|_inf; first right-hand side
                uint32_t qm = 0u;
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    const int cb8 = (chunk0 + ch) * 8;
                    const int jb = row_base + ch * 8;
                    float r[8];
                    if (state_chunk_owner && ch == kChunks - 1) {
                        // contraction slots 152, 153, 154 = axis state relative to the origin
#pragma unroll
                        for (int e = 0; e < 8; ++e) r[e] = 0.f;
                        r[1] = live ? x1f : 0.f;
                        r[2] = live ? x2f : 0.f;
                    } else {
                        uint32_t tv[8], lov[8];
                        tmem_ld8(tmem + kColD + cb8, tv);
                        tmem_ld8(tmem + kColLo + cb8, lov);
                        tmem_wait_ld();
                        tmem_st8(tmem + kColT + cb8, tv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int j = jb + e;
                            const float t = __uint_as_float(tv[e]), l = __uint_as_float(lov[e]), h = hi[ch * 8 + e];
                            const float qc = P.qhalf * (l + h);
                            const float qv = -qc - fmaf(S.gpx[j * 3 + 2], x2f, S.gpx[j * 3 + 1] * x1f);
                            qm = max(qm, __float_as_uint(fabsf(qv)));
                            const float zc = clipf(t, l, h);
                            r[e] = fmaf(P.rho, 2.f * zc - t, qc);
                        }
                    }
                    publish_rhs_pair(sAh, sAl, jb >> 3, s, r);
                }
                tmem_wait_st();
                S.Qm[hf * kParts + part][s] = qm;
                proxy_fence();
            }
            bool done = !live;
            bool failed = false;               // stopped without converging (residual stalled far above the threshold)
            uint32_t checkpoint = 0xffffffffu; // residual 32 iterations ago
            int iters = 0;
            tc_fence_before();
            int local_done = __syncthreads_and(done ? 1 : 0);      // right-hand sides published, |q| partials visible
            tc_fence_after();
            {
                uint32_t m = S.Qm[0][s];
#pragma unroll
                for (int p4 = 1; p4 < 2 * kParts; ++p4) m = max(m, S.Qm[p4][s]);
                thr = P.eps_stop * fmaxf(1.f, __uint_as_float(m));
            }

            PC_MARK(pc_pro);
            // ================= ADMM =================
            for (int k = 0; k < P.max_iter; ++k) {
                // ---- the pair meets: "my right-hand sides are published; all my columns are done / not done" ----
                ++exchange;
                if (tid == 0) remote_store_release(other_word_addr, (exchange << 1) | (uint32_t)local_done);
                uint32_t word;
                do { word = load_acquire_cluster(&S.other_word); } while ((word >> 1) < exchange);
                // (a word from a LATER exchange means the other CTA went on: the pair was not done at this one)
                const bool pair_done = local_done && (word >> 1) == exchange && (word & 1u);
                PC_MARK(pc_exch);
                if (pair_done) break;
                if (rank == 0 && warp == 0) {
                    // the whole warp runs this (warp-uniform operands); one elected lane issues
                    tc_fence_after();
                    const uint64_t a_hi = operand_desc(smem_u32(sAh), kLboAp), a_lo = operand_desc(smem_u32(sAl), kLboAp);
                    const uint64_t b_hi = operand_desc(smem_u32(sBh), kLboBp), b_lo = operand_desc(smem_u32(sBl), kLboBp);
                    const uint32_t d = tmem_d;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t a = pass == 1 ? a_lo : a_hi, bb = pass == 2 ? b_lo : b_hi;
#pragma unroll
                        for (int ks = 0; ks < kF16Tile / 16; ++ks)      // K chunk pairs: the start-address field advances by 2 LBO / 16
                            umma_f16_pair(d, a + (uint64_t)(ks * 2 * kLboAp / 16), bb + (uint64_t)(ks * 2 * kLboBp / 16), idesc,
                                          (pass | ks) ? 1u : 0u);
                    }
                    umma_commit_pair(&S.bar_mma);
                    __syncwarp();
                }
                {
                    mbar_wait(&S.bar_mma, mma_parity);
                    tc_fence_after();
                    PC_MARK(pc_mma);
                    uint32_t rd = 0u;
                    // tcgen05.ld / st are warp-collective: the warp runs the epilogue while any of its
                    // columns is still iterating; converged columns keep t (and republish the same r)
                    if (__any_sync(0xffffffffu, !done)) {
                        const float relax = done ? 0.f : P.alpha;
#pragma unroll
                        for (int ch = 0; ch < kChunks; ++ch) {
                            if (state_chunk_owner && ch == kChunks - 1) continue;
                            const int cb8 = (chunk0 + ch) * 8;
                            uint32_t vd[8], vt[8], vl[8];
                            tmem_ld8(tmem + kColD + cb8, vd);
                            tmem_ld8(tmem + kColT + cb8, vt);
                            tmem_ld8(tmem + kColLo + cb8, vl);
                            tmem_wait_ld();
                            float r[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float t = __uint_as_float(vt[e]), l = __uint_as_float(vl[e]), h = hi[ch * 8 + e];
                                const float zo = clipf(t, l, h);
                                const float dd = __uint_as_float(vd[e]) - zo;
                                const float tn = fmaf(relax, dd, t);        // alpha v + (1 - alpha) zo + (t - zo)
                                const float zn = clipf(tn, l, h);
                                rd = max(rd, __float_as_uint(fabsf(dd)));
                                vt[e] = __float_as_uint(tn);
                                r[e] = fmaf(P.qhalf, l + h, fmaf(P.two_rho, zn, -P.rho * tn));
                            }
                            tmem_st8(tmem + kColT + cb8, vt);
                            publish_rhs_pair(sAh, sAl, (row_base >> 3) + ch, s, r);
                        }
                        tmem_wait_st();
                        rd = done ? 0u : rd;
                    }
                    S.Res[hf * kParts + part][s] = rd;
                    proxy_fence();
                }
                mma_parity ^= 1u;
                tc_fence_before();
                __syncthreads();          // new right-hand sides and residual partials are visible; D is consumed
                if (!done) {
                    uint32_t m = S.Res[0][s];
#pragma unroll
                    for (int p4 = 1; p4 < 2 * kParts; ++p4) m = max(m, S.Res[p4][s]);
                    ++iters;
                    done = __uint_as_float(m) <= thr;
                    // A column whose residual has not halved over 32 iterations while still 16x above its
                    // threshold is not going to make it (a diverged walk: FP32 noise on a huge linear term)
                    if (!done && (k & 31) == 31) {
                        if (checkpoint != 0xffffffffu && m + (1u << 23) > checkpoint && __uint_as_float(m) > 16.f * thr) {
                            done = true;
                            failed = true;
                        }
                        checkpoint = m;
                    }
                }
                ++unit_iters;
                local_done = __syncthreads_and(done ? 1 : 0);     // partials consumed before the next epilogue rewrites them
                tc_fence_after();
                PC_MARK(pc_epi);
            }

            // ================= apply the first sample (FP64) =================
            if (B.zmp_plan) {
                const bool emit = live && i == nticks - 1;
                double* pz = B.zmp_plan + (long long)(emit ? c : 0) * N;
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    if (state_chunk_owner && ch == kChunks - 1) continue;
                    const int cb8 = (chunk0 + ch) * 8;
                    const int jb = row_base + ch * 8;
                    uint32_t vt[8], vl[8];
                    tmem_ld8(tmem + kColT + cb8, vt);
                    tmem_ld8(tmem + kColLo + cb8, vl);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = jb + e;
                        if (emit && j < N) {
                            int smp = i + B.bound_index_bias + j;
                            smp = smp > last ? last : smp;
                            const double l64 = __ldg(plo + (long long)smp * B.bound_stride);
                            const double h64 = __ldg(phi + (long long)smp * B.bound_stride);
                            const double z = origin + (double)clipf(__uint_as_float(vt[e]), __uint_as_float(vl[e]), hi[ch * 8 + e]);
                            pz[j] = fmin(fmax(z, l64), h64);
                        }
                    }
                }
            }
            uint32_t vt0[8], vl0[8];
            if (hf == 0 && part == 0) {        // warp-collective loads first, then the columns that are live
                tmem_ld8(tmem + kColT, vt0);
                tmem_ld8(tmem + kColLo, vl0);
                tmem_wait_ld();
            }
            if (head && live) {
                const float t0 = __uint_as_float(vt0[0]), l0f = __uint_as_float(vl0[0]), h0f = hi[0];
                int smp = i + B.bound_index_bias;
                smp = smp > last ? last : smp;
                const double l0 = __ldg(plo + (long long)smp * B.bound_stride);
                const double h0 = __ldg(phi + (long long)smp * B.bound_stride);
                const bool act_lo = t0 <= l0f, act_hi = t0 >= h0f;
                double z0 = act_lo ? l0 : (act_hi ? h0 : origin + (double)t0);
                z0 = fmin(fmax(z0, l0), h0);          // the applied sample is feasible in FP64
                const double x0 = S.X[s][0], x1 = S.X[s][1], x2 = S.X[s][2];
                const double pred = fma(P.px02, x2, fma(P.T, x1, x0));
                const double jerk = (z0 - pred) / P.pu00;
                const double n0 = fma(P.c3, jerk, fma(P.c2, x2, fma(P.T, x1, x0)));
                double n1 = fma(P.c2, jerk, fma(P.T, x2, x1));
                const double n2 = fma(P.T, jerk, x2);
                const double zmp = n0 - P.hg * n2;
                const double below = l0 - zmp, above = zmp - h0;
                viol = fmax(viol, fmax(fmax(below, above), 0.0));
                lean = fmax(lean, fabs(n0 - zmp));
                unsigned f = (below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) | (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u) |
                             (act_lo ? LIPM_FLAG_ACTIVE_LO : 0u) | (act_hi ? LIPM_FLAG_ACTIVE_HI : 0u);
                if (!done || failed) { f |= LIPM_FLAG_NOT_CONVERGED; notconv += 1; }
                if (i == ftick) n1 -= imp;
                const long long sample = i + 1;
                if (ph) {
                    double* dst = ph + (sample - B.hist_index_bias) * 3;
                    dst[0] = n0; dst[1] = n1; dst[2] = n2;
                }
                if (pf) pf[sample * 2] = (uint8_t)f;
                if (pcom) pcom[sample * 2] = n0;
                if (pzmp) pzmp[sample * 2] = zmp;
                S.X[s][0] = n0; S.X[s][1] = n1; S.X[s][2] = n2;
                iter_sum += iters;
            }
            // the reference raises at the first QP without a solution (zmp_controller.py:193-194): the walk of
            // this column ends here (every thread of the column reaches the same verdict from the same residuals)
            if (live && (!done || failed)) nticks = i + 1;
            tc_fence_before();
            __syncthreads();              // the new state is visible to every thread of the column
            tc_fence_after();
            PC_MARK(pc_apply);
        }
#ifdef LIPM_F16P_PROFILE
        if (tid == 0 && blockIdx.x < 2)
            printf("cta %d unit %u ticks %d group-iters %u: prologue %lld exchange %lld mma %lld epilogue %lld apply %lld (cycles)\n", (int)blockIdx.x, unit,
                   max_ticks, unit_iters, pc_pro, pc_exch, pc_mma, pc_epi, pc_apply);
#endif

        // ---- per-column summaries ---------------------------------------------------------------
        if (tid == 0 && rank == 0) {          // one count per 128-column unit
            atomicAdd(P.stats, (unsigned long long)unit_iters);
            atomicAdd(P.stats + 1, 1ull);
        }
        if (head && valid) {
            if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
            if (B.lean_max) B.lean_max[b * 2 + axis] = lean;
            if (B.status) B.status[b * 2 + axis] = notconv;
            if (B.iters) B.iters[b * 2 + axis] = iter_sum;
        }
    }

    tc_fence_before();
    cluster_sync_all();                   // neither CTA frees tensor memory while the other may still be addressed by an MMA
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(S.tmem_base) : "memory");
}

template <int kParts>
int launch_pair(lipm_plan* plan, PairParams& P, cudaStream_t stream, int& pairs_out) {
    auto kernel = k_strict_rollout_f16p<kParts>;
    const size_t smem = kSmemOperandsP + sizeof(PairControl<kParts>);
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need the whole shared-memory carve-out (the default split follows ONE block's request)
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    // pairs that can be resident at once (a pair = one cluster of two CTAs on the two SMs of a TPC)
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(128 * kParts);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(2 * (unsigned)plan->sm_count);
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, kernel, &cfg) != cudaSuccess || resident <= 0) {
        cudaGetLastError();
        resident = plan->sm_count / 2;
    }
    // The occupancy calculator counts one CTA per SM for any kernel that allocates tensor memory; the hardware
    // co-schedules two 256-column CTAs (measured), so the two-per-SM variant asks for twice the pairs.
    if (kParts == 2) resident *= 2;
    const unsigned pairs = P.num_units < (unsigned)resident ? P.num_units : (unsigned)resident;
    pairs_out = (int)pairs;
#ifdef LIPM_F16P_PROFILE
    {
        int blocks = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, 128 * kParts, smem);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kernel);
        printf("blocks per SM %d, regs %d, static smem %zu, local %zu maxthreads %d maxdyn %d carve %d\n", blocks, fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout);
        for (size_t sm = 32768; sm <= 114688; sm += 8192) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, 128 * kParts, sm); printf("  smem %zu -> %d blocks\n", sm, blocks); }
    }
    printf("pair kernel: parts %d, %u units, resident clusters %d, launching %u pairs, smem %zu\n", kParts, P.num_units, resident, pairs, smem);
#endif
    cfg.gridDim = dim3(2 * pairs);
    LIPM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, P));
    return LIPM_OK;
}

}  // namespace

int launch_strict_rollout_f16p(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (!plan->d_bf16 || plan->N > kMaxReducedHorizon) {
        set_error("reduced-precision strict mode needs a plan created with precision F16X3 and horizon <= %d", kMaxReducedHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    PairParams P;
    P.cb = cb;
    P.N = plan->N;
    P.max_iter = d.max_iter;
    P.rho = (float)d.rho; P.two_rho = (float)(2.0 * d.rho); P.alpha = (float)d.alpha; P.qhalf = (float)(0.5 * d.Q);
    P.eps_stop = (float)(d.eps / std::fmax(1.0 + d.alpha, d.rho * d.alpha));
    P.viol_tol = d.viol_tol;
    P.T = plan->T; P.c2 = std::pow(d.dt, 2) / 2.0; P.c3 = std::pow(d.dt, 3) / 6.0; P.hg = plan->hg;
    P.px02 = P.c2 - P.hg;
    P.pu00 = plan->pu00;
    P.bimg = plan->d_bf16;
    P.gpx = plan->d_gpx;
    P.num_units = (unsigned)((cb.ncols + kF16Cols - 1) / kF16Cols);
    if (P.num_units == 0) return LIPM_OK;
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    P.tile_counter = state.counter();
    P.stats = state.stats();
    int pairs = 0;
    // few units: one 20-warp CTA per SM (a column's rows spread over 10 threads, every unit on its own pair of SMs);
    // many units: two 8-warp CTAs per SM (one's epilogue overlaps the other's contraction)
    const int rc = (2 * (int)P.num_units <= plan->sm_count) ? launch_pair<5>(plan, P, stream, pairs) : launch_pair<2>(plan, P, stream, pairs);
    if (rc != LIPM_OK) return rc;
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(state.publish(plan->d_stats));
    return LIPM_OK;
}

}  // namespace lipm
