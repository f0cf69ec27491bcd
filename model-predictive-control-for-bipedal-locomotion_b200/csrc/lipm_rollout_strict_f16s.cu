// Strict (box-constrained) rollout, reduced-precision path for LONG horizons (152 < N <= 637):
// the operands of the ADMM contraction no longer fit in shared memory, so they are STREAMED.
//
// Same algorithm, outputs, precision scheme and reference code as lipm_rollout_strict_f16.cu
// (zmp_controller.py:173-195 strict QP, :199 state update, :81-106 windowing / padding / push;
// BASELINE config 3: horizon 300 / 600): v = M r as three FP16 hi/lo passes with FP32 accumulation
// in tensor memory, FP32 ADMM state in tick-relative coordinates, FP64 walk state and outputs.
//
// Organisation (one persistent CTA per SM, 128 columns = 128 tensor-memory lanes per CTA):
//   * the iteration matrix is cut into SLABS of 160 output rows; D_slab[column, row] =
//     sum_k A[column, k] B_slab[row, k] with A = the right-hand sides of the CTA's 128 columns and
//     B = [M | M G Px] (the axis state rides in contraction slots N .. N+2, as in the resident kernel);
//   * both operands live in global memory (L2) as ready-made K-major operand images, FP16 hi and lo:
//     B once per plan, A in a per-CTA workspace that the epilogue writes.  A COPY thread moves them
//     in 32-slot stages (8 + 8 + 10 + 10 KB) with cp.async.bulk through a 5-deep mbarrier ring;
//   * an MMA thread issues, per stage, 2 x 3 tcgen05.mma (M = 128, N = 160, K = 16) and commits the
//     stage back to the ring; a slab's accumulator is one of TWO 160-column tensor-memory buffers, so
//     the contraction of slab s+1 runs while the 16 epilogue warps work on slab s;
//   * epilogue thread (quadrant q = warp % 4, part = warp / 4) owns column 32 q + lane and rows
//     40 part .. 40 part + 39 of every slab.  Its ADMM state (t, lo, hi; FP32) is in the CTA's
//     workspace, [row][column] so that a warp reads 128 contiguous bytes per row; it applies
//     relaxation, box projection, dual update and residual and writes the next right-hand side,
//     split hi / lo, straight into the other A image (16-byte stores, 512 contiguous bytes per warp);
//   * right-hand sides are double buffered (an iteration reads one image and writes the other), so a
//     column that has converged republishes its frozen right-hand side twice and then idles.
// Results do not depend on which columns share a CTA (per-column stop test, per-column freeze).
#include "lipm_internal.cuh"
#include "lipm_tcgen05.cuh"

#include <cuda_fp16.h>

namespace lipm {

namespace {

using namespace tc;

constexpr int kSlab = kF16Tile;                      // 160 output rows per accumulator buffer = MMA N
constexpr int kEpiWarps = 16;
constexpr int kParts = 4;
constexpr int kPartRows = 40;
constexpr int kChunks = 5;                           // 8-row chunks per part and slab
constexpr int kCopyWarp = kEpiWarps;                 // warp 16: lane 0 streams the operands
constexpr int kMmaWarp = kEpiWarps + 1;              // warp 17: lane 0 issues the MMAs
constexpr int kThreadsS = 32 * (kEpiWarps + 2);      // 576
constexpr int kStageGroups = 4;                      // 8-slot contraction groups per stage (K = 32)
constexpr int kStages = 5;
constexpr int kGroupA = (kF16Cols / 8) * 128;        // 2048 B: 8 contraction slots of 128 columns (= LBO of A)
constexpr int kGroupB = (kSlab / 8) * 128;           // 2560 B: 8 contraction slots of 160 rows    (= LBO of B)
constexpr int kStageA = kStageGroups * kGroupA;      // 8192
constexpr int kStageB = kStageGroups * kGroupB;      // 10240
constexpr int kStageBytes = 2 * kStageA + 2 * kStageB;   // 36864: A hi, A lo, B hi, B lo

struct StreamParams {
    ColumnBatch cb;
    int N, NK, slabs;            // horizon, padded contraction length (multiple of 16), 160-row slabs
    int max_iter;
    float rho, two_rho, alpha, qhalf;
    float eps_stop;
    double viol_tol;
    double T, c2, c3, hg, px02, pu00;
    const unsigned char* bimg;   // hi image, then lo image: [slab][NK/8][160 rows][16 B]
    size_t bimg_bytes;           // one image
    const double* gpx;           // [padded8(N),3]
    unsigned char* work;         // per-CTA workspace
    size_t work_stride;
    unsigned int* tile_counter;
    unsigned long long* stats;
    unsigned int num_units;
};

struct ControlS {
    unsigned long long full[kStages], empty[kStages], dfull[2], dfree[2];
    uint32_t tmem_base;
    int unit, max_ticks, pad;
    float gpx[kMaxReducedStreamK * 3];
    double X[kF16Cols][3];
    uint32_t Res[kParts][kF16Cols];
    uint32_t Qm[kParts][kF16Cols];
};

constexpr size_t kSmemS = (size_t)kStages * kStageBytes + sizeof(ControlS);

// Eight consecutive contraction slots of one column (8-slot group kg), split into FP16 hi + lo, as two
// 16-byte stores into the operand images in global memory.
__device__ __forceinline__ void publish_rhs(unsigned char* Rh, unsigned char* Rl, int kg, int s, const float (&r)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(r[2 * e], r[2 * e + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(r[2 * e] - back.x, r[2 * e + 1] - back.y);
        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const size_t off = (size_t)kg * kGroupA + s * 16;
    *reinterpret_cast<uint4*>(Rh + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(Rl + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kThreadsS, 1) k_strict_rollout_f16s(const StreamParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    ControlS& S = *reinterpret_cast<ControlS*>(smem + (size_t)kStages * kStageBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool epi = warp < kEpiWarps;
    const int quad = warp & 3, part = epi ? (warp >> 2) : kParts;
    const int s = quad * 32 + lane;                  // column = tensor-memory lane (epilogue warps)
    const int N = P.N, NK = P.NK, slabs = P.slabs;
    const int groups = NK / 8;                       // 8-slot contraction groups
    const int stages_per_slab = (groups + kStageGroups - 1) / kStageGroups;
    const ColumnBatch& B = P.cb;

    // workspace of this CTA: t (two copies, alternating per tick), lo, hi as [row][128] FP32, then the
    // right-hand-side operand images [buffer][hi, lo][groups][128 columns][16 B]
    unsigned char* W = P.work + (size_t)blockIdx.x * P.work_stride;
    float* Tb0 = reinterpret_cast<float*>(W);
    float* Tb1 = Tb0 + (size_t)NK * kF16Cols;
    float* LO = Tb1 + (size_t)NK * kF16Cols;
    float* HI = LO + (size_t)NK * kF16Cols;
    unsigned char* Rimg = reinterpret_cast<unsigned char*>(HI + (size_t)NK * kF16Cols);
    const size_t rimg_bytes = (size_t)groups * kGroupA;           // one image
    const unsigned char* Bh = P.bimg;
    const unsigned char* Bl = P.bimg + P.bimg_bytes;

    // ---- one-time setup: barriers, tensor memory, G Px -------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&S.dfull[i], 1); mbar_init(&S.dfree[i], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < NK * 3; i += kThreadsS) S.gpx[i] = (i < N * 3) ? (float)P.gpx[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base + ((uint32_t)(32 * quad) << 16);   // this warp's lane quadrant
    // D = FP32, A = B = FP16, both K-major, N = 160, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kSlab >> 3) << 17) | ((uint32_t)(kF16Cols >> 4) << 24);

    uint32_t fill = 0;        // copy thread: stages requested so far
    uint32_t cons = 0;        // MMA thread: stages consumed so far
    uint32_t dcount = 0;      // MMA thread and epilogue warps: slabs so far (accumulator buffer = dcount & 1)
    int rbuf = 0;             // right-hand-side image the next contraction reads

    for (;;) {
        // ---- fetch a 128-column unit ---------------------------------------------------------------
        __syncthreads();
        if (tid == 0) {
            S.unit = (int)atomicAdd(P.tile_counter, 1u);
            S.max_ticks = 0;
        }
        __syncthreads();
        const unsigned int unit = (unsigned int)S.unit;
        if (unit >= P.num_units) break;
        const int c = (B.unit_order ? B.unit_order[unit] : (int)(P.num_units - 1u - unit)) * kF16Cols + s;

        const bool valid = epi && c < B.ncols;
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
        if (part == 0) {
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
        const int max_ticks = S.max_ticks;
        unsigned int unit_iters = 0;
        double origin = 0.0;          // origin of the tick-relative coordinates (FP64)

        for (int i = 0; i < max_ticks; ++i) {
            const bool live = i < nticks;
            float* Tcur = (i & 1) ? Tb1 : Tb0;
            const float* Told = (i & 1) ? Tb0 : Tb1;
            float thr = 0.f;
            float x1f = 0.f, x2f = 0.f;
            if (epi) {
                // ================= tick prologue: window, warm start, |q|_inf, first right-hand side =================
                const double x0d = S.X[s][0];
                x1f = live ? (float)S.X[s][1] : 0.f;
                x2f = live ? (float)S.X[s][2] : 0.f;
                const float shift = (float)(origin - x0d);     // re-reference the warm start
                origin = x0d;
                unsigned char* Rh = Rimg + (size_t)rbuf * 2 * rimg_bytes;
                unsigned char* Rl = Rh + rimg_bytes;
                uint32_t qm = 0u;
                for (int slab = 0; slab < slabs; ++slab) {
#pragma unroll 1
                    for (int ch = 0; ch < kChunks; ++ch) {
                        const int jb = slab * kSlab + part * kPartRows + ch * 8;
                        if (jb >= NK) continue;
                        float r[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int j = jb + e;
                            float l = 0.f, h = 0.f, t = 0.f, rv = 0.f;
                            if (live && j < N) {
                                int smp = i + B.bound_index_bias + j;
                                smp = smp > last ? last : smp;
                                l = (float)(__ldg(plo + (long long)smp * B.bound_stride) - x0d);
                                h = (float)(__ldg(phi + (long long)smp * B.bound_stride) - x0d);
                                if (i == 0) t = 0.5f * (l + h);
                                else if (j + 1 < N) t = Told[(size_t)(j + 1) * kF16Cols + s] + shift;
                                else t = clipf(Told[(size_t)j * kF16Cols + s] + shift, l, h);   // last row: keep the primal value, drop the dual
                                const float qc = P.qhalf * (l + h);
                                const float qv = -qc - fmaf(S.gpx[j * 3 + 2], x2f, S.gpx[j * 3 + 1] * x1f);
                                qm = max(qm, __float_as_uint(fabsf(qv)));
                                const float zc = clipf(t, l, h);
                                rv = fmaf(P.rho, 2.f * zc - t, qc);
                            }
                            if (j == N + 1) rv = x1f;          // contraction slots N, N+1, N+2 = axis state relative to the origin
                            if (j == N + 2) rv = x2f;
                            __stcs(Tcur + (size_t)j * kF16Cols + s, t);
                            __stcs(LO + (size_t)j * kF16Cols + s, l);
                            __stcs(HI + (size_t)j * kF16Cols + s, h);
                            r[e] = rv;
                        }
                        publish_rhs(Rh, Rl, jb >> 3, s, r);
                    }
                }
                S.Qm[part][s] = qm;
                proxy_fence_global();
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            bool done = !live;
            bool failed = false;               // stopped without converging (residual stalled far above the threshold)
            uint32_t checkpoint = 0xffffffffu; // residual 32 iterations ago
            int iters = 0;
            int age = 0;                       // iterations since this column froze (it republishes its right-hand side twice)
            if (epi) {
                uint32_t m = S.Qm[0][s];
#pragma unroll
                for (int p4 = 1; p4 < kParts; ++p4) m = max(m, S.Qm[p4][s]);
                thr = P.eps_stop * fmaxf(1.f, __uint_as_float(m));
            }

            // ================= ADMM =================
            for (int k = 0; k < P.max_iter; ++k) {
                const unsigned char* RhC = Rimg + (size_t)rbuf * 2 * rimg_bytes;      // read by this contraction
                unsigned char* RhN = Rimg + (size_t)(rbuf ^ 1) * 2 * rimg_bytes;      // written by this epilogue
                if (warp == kCopyWarp) {
                    if (lane == 0) {
                        for (int slab = 0; slab < slabs; ++slab) {
                            for (int st = 0; st < stages_per_slab; ++st) {
                                const uint32_t slot = fill % kStages, phase = (fill / kStages) & 1u;
                                mbar_wait(&S.empty[slot], phase ^ 1u);
                                const int g0 = st * kStageGroups;
                                const int ng = (groups - g0 < kStageGroups) ? (groups - g0) : kStageGroups;
                                const uint32_t bytesA = (uint32_t)ng * kGroupA, bytesB = (uint32_t)ng * kGroupB;
                                mbar_expect_tx(&S.full[slot], 2u * bytesA + 2u * bytesB);
                                unsigned char* dst = smem + (size_t)slot * kStageBytes;
                                bulk_load(dst, RhC + (size_t)g0 * kGroupA, bytesA, &S.full[slot]);
                                bulk_load(dst + kStageA, RhC + rimg_bytes + (size_t)g0 * kGroupA, bytesA, &S.full[slot]);
                                const size_t boff = ((size_t)slab * groups + g0) * kGroupB;
                                bulk_load(dst + 2 * kStageA, Bh + boff, bytesB, &S.full[slot]);
                                bulk_load(dst + 2 * kStageA + kStageB, Bl + boff, bytesB, &S.full[slot]);
                                ++fill;
                            }
                        }
                    }
                    __syncwarp();
                } else if (warp == kMmaWarp) {
                    if (lane == 0) {
                        for (int slab = 0; slab < slabs; ++slab) {
                            const uint32_t buf = dcount & 1u, use = dcount >> 1;
                            mbar_wait(&S.dfree[buf], (use & 1u) ^ 1u);      // the epilogue has drained this accumulator
                            tc_fence_after();
                            const uint32_t d = S.tmem_base + buf * kSlab;
                            for (int st = 0; st < stages_per_slab; ++st) {
                                const uint32_t slot = cons % kStages, phase = (cons / kStages) & 1u;
                                mbar_wait(&S.full[slot], phase);
                                tc_fence_after();
                                const int g0 = st * kStageGroups;
                                const int ng = (groups - g0 < kStageGroups) ? (groups - g0) : kStageGroups;
                                const uint32_t base = smem_u32(smem + (size_t)slot * kStageBytes);
                                const uint32_t aH = base, aL = base + kStageA, bH = base + 2 * kStageA, bL = bH + kStageB;
                                for (int ks = 0; ks < ng / 2; ++ks) {
#pragma unroll
                                    for (int pass = 0; pass < 3; ++pass) {
                                        const uint32_t a = pass == 1 ? aL : aH, bb = pass == 2 ? bL : bH;
                                        umma_f16(d, operand_desc(a + ks * 2 * kGroupA, kGroupA), operand_desc(bb + ks * 2 * kGroupB, kGroupB),
                                                 idesc, (st | ks | pass) ? 1u : 0u);
                                    }
                                }
                                umma_commit(&S.empty[slot]);        // the stage is free once these MMAs have read it
                                ++cons;
                            }
                            umma_commit(&S.dfull[buf]);
                            ++dcount;
                        }
                    }
                    __syncwarp();
                } else {
                    uint32_t rd = 0u;
                    const float relax = done ? 0.f : P.alpha;      // a converged column keeps its t and republishes the same r
                    const bool work = __any_sync(0xffffffffu, age < 2);
                    unsigned char* RlN = RhN + rimg_bytes;
                    // chunks of 8 rows, (slab, chunk) in order; the state of the NEXT chunk is fetched while this one
                    // is computed (the loads come from L2 / HBM: one chunk of latency is hidden per chunk)
                    const int nitems = slabs * kChunks;
                    float tn[8], ln[8], hn[8];
                    {
                        const int jn = part * kPartRows;
                        if (work && jn < NK) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const size_t at = (size_t)(jn + e) * kF16Cols + s;
                                tn[e] = __ldcs(Tcur + at); ln[e] = __ldcs(LO + at); hn[e] = __ldcs(HI + at);
                            }
                        }
                    }
#pragma unroll 1
                    for (int idx = 0; idx < nitems; ++idx) {
                        const int slab = idx / kChunks, ch = idx - slab * kChunks;
                        const int jb = slab * kSlab + part * kPartRows + ch * 8;
                        const uint32_t buf = dcount & 1u, use = dcount >> 1;
                        float t[8], l[8], h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { t[e] = tn[e]; l[e] = ln[e]; h[e] = hn[e]; }
                        if (idx + 1 < nitems) {
                            const int sn = (idx + 1) / kChunks, cn = (idx + 1) - sn * kChunks;
                            const int jn = sn * kSlab + part * kPartRows + cn * 8;
                            if (work && jn < NK) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const size_t at = (size_t)(jn + e) * kF16Cols + s;
                                    tn[e] = __ldcs(Tcur + at); ln[e] = __ldcs(LO + at); hn[e] = __ldcs(HI + at);
                                }
                            }
                        }
                        if (ch == 0) {
                            mbar_wait(&S.dfull[buf], use & 1u);
                            tc_fence_after();
                        }
                        if (work && jb < NK) {
                            uint32_t vd[8];
                            tmem_ld8(tmem + buf * kSlab + part * kPartRows + ch * 8, vd);
                            tmem_wait_ld();
                            float r[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int j = jb + e;
                                const float zo = clipf(t[e], l[e], h[e]);
                                const float dd = __uint_as_float(vd[e]) - zo;
                                const float tnew = fmaf(relax, dd, t[e]);        // alpha v + (1 - alpha) zo + (t - zo)
                                const float zn = clipf(tnew, l[e], h[e]);
                                float rv = fmaf(P.qhalf, l[e] + h[e], fmaf(P.two_rho, zn, -P.rho * tnew));
                                if (j < N) {
                                    rd = max(rd, __float_as_uint(fabsf(dd)));
                                    __stcs(Tcur + (size_t)j * kF16Cols + s, tnew);
                                } else {
                                    rv = (j == N + 1) ? x1f : ((j == N + 2) ? x2f : 0.f);
                                }
                                r[e] = rv;
                            }
                            publish_rhs(RhN, RlN, jb >> 3, s, r);
                        }
                        if (ch == kChunks - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&S.dfree[buf]);
                            ++dcount;
                        }
                    }
                    S.Res[part][s] = done ? 0u : rd;
                    proxy_fence_global();
                }
                tc_fence_before();
                __syncthreads();          // new right-hand sides and residual partials are visible
                bool mine = true;
                if (epi) {
                    if (done) {
                        ++age;
                    } else {
                        uint32_t m = S.Res[0][s];
#pragma unroll
                        for (int p4 = 1; p4 < kParts; ++p4) m = max(m, S.Res[p4][s]);
                        ++iters;
                        done = __uint_as_float(m) <= thr;
                        // same give-up rule as the resident kernel: residual not halved over 32 iterations while 16x above the threshold
                        if (!done && (k & 31) == 31) {
                            if (checkpoint != 0xffffffffu && m + (1u << 23) > checkpoint && __uint_as_float(m) > 16.f * thr) {
                                done = true;
                                failed = true;
                            }
                            checkpoint = m;
                        }
                    }
                    mine = done;
                }
                ++unit_iters;
                rbuf ^= 1;
                const int stop = __syncthreads_and(mine ? 1 : 0);
                tc_fence_after();
                if (stop) break;
            }

            // ================= apply the first sample (FP64) =================
            if (epi && B.zmp_plan) {
                const bool emit = live && i == nticks - 1;
                double* pz = B.zmp_plan + (long long)(emit ? c : 0) * N;
                for (int slab = 0; slab < slabs; ++slab) {
                    for (int ch = 0; ch < kChunks; ++ch) {
                        const int jb = slab * kSlab + part * kPartRows + ch * 8;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int j = jb + e;
                            if (emit && j < N) {
                                int smp = i + B.bound_index_bias + j;
                                smp = smp > last ? last : smp;
                                const double l64 = __ldg(plo + (long long)smp * B.bound_stride);
                                const double h64 = __ldg(phi + (long long)smp * B.bound_stride);
                                const size_t at = (size_t)j * kF16Cols + s;
                                const double z = origin + (double)clipf(Tcur[at], LO[at], HI[at]);
                                pz[j] = fmin(fmax(z, l64), h64);
                            }
                        }
                    }
                }
            }
            if (part == 0 && live) {
                const float t0 = Tcur[s], l0f = LO[s], h0f = HI[s];
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
            // this column ends here (every part of the column reaches the same verdict from the same residuals)
            if (live && (!done || failed)) nticks = i + 1;
            __syncthreads();              // the new state and the final t are visible to every part
        }

        // ---- per-column summaries ---------------------------------------------------------------
        if (tid == 0) {
            atomicAdd(P.stats, (unsigned long long)unit_iters);
            atomicAdd(P.stats + 1, 1ull);
        }
        if (part == 0 && valid) {
            if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
            if (B.lean_max) B.lean_max[b * 2 + axis] = lean;
            if (B.status) B.status[b * 2 + axis] = notconv;
            if (B.iters) B.iters[b * 2 + axis] = iter_sum;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(S.tmem_base) : "memory");
}

}  // namespace

int launch_strict_rollout_f16s(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (!plan->d_f16s) {
        set_error("reduced-precision strict mode for horizon %d needs a plan created with precision F16X3", plan->N);
        return LIPM_E_UNSUPPORTED;
    }
    StreamParams P;
    P.cb = cb;
    P.N = plan->N; P.NK = plan->f16s_nk; P.slabs = plan->f16s_slabs;
    P.max_iter = d.max_iter;
    P.rho = (float)d.rho; P.two_rho = (float)(2.0 * d.rho); P.alpha = (float)d.alpha; P.qhalf = (float)(0.5 * d.Q);
    P.eps_stop = (float)(d.eps / std::fmax(1.0 + d.alpha, d.rho * d.alpha));
    P.viol_tol = d.viol_tol;
    P.T = plan->T; P.c2 = std::pow(d.dt, 2) / 2.0; P.c3 = std::pow(d.dt, 3) / 6.0; P.hg = plan->hg;
    P.px02 = P.c2 - P.hg;
    P.pu00 = plan->pu00;
    P.bimg = plan->d_f16s;
    P.bimg_bytes = (size_t)P.slabs * (P.NK / 8) * kGroupB;
    P.gpx = plan->d_gpx;
    P.num_units = (unsigned)((cb.ncols + kF16Cols - 1) / kF16Cols);
    if (P.num_units == 0) return LIPM_OK;
    LIPM_CUDA_TRY(cudaFuncSetAttribute(k_strict_rollout_f16s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemS));
    const unsigned grid = P.num_units < (unsigned)plan->sm_count ? P.num_units : (unsigned)plan->sm_count;
    // per-CTA workspace: 4 FP32 row images (t, t, lo, hi) + 2 x (hi, lo) right-hand-side operand images
    P.work_stride = (size_t)P.NK * kF16Cols * 4 * sizeof(float) + 4 * (size_t)(P.NK / 8) * kGroupA;
    StreamBuffer work;
    LIPM_CUDA_TRY(work.alloc((size_t)grid * P.work_stride, stream));
    P.work = work.as<unsigned char>();
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    P.tile_counter = state.counter();
    P.stats = state.stats();
    k_strict_rollout_f16s<<<grid, kThreadsS, kSmemS, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(state.publish(plan->d_stats));
    return LIPM_OK;
}

}  // namespace lipm
