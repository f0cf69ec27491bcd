// Strict (box-constrained) receding-horizon rollout, reduced-precision path: the ADMM contraction on
// the 5th-generation tensor cores (tcgen05.mma, FP16 operands, FP32 accumulators in tensor memory).
//
// Same algorithm, same outputs and same reference code as lipm_rollout_strict.cu
// (zmp_controller.py:173-195 strict QP, :199 state update, :81-106 windowing / padding / push);
// only the arithmetic of the iteration differs.  north_star: trajectories within 1e-4 m.
//
// Precision scheme (tools/precision_probe.py is the CPU emulation that picked it):
//   * v = M r is evaluated as  M_hi r_hi + M_hi r_lo + M_lo r_hi  with FP16 hi/lo splits of both
//     operands (22 significant bits) and FP32 accumulation: three kind::f16 passes over K = 160.
//     A single TF32 or BF16 pass does not even converge (the rounding noise is amplified by the
//     fixed-point iteration); BF16 x3 reaches 2e-5 m, FP16 x3 reaches 1e-6 m.
//   * the ADMM state (t, lo, hi) is FP32 in TICK-RELATIVE coordinates (origin = the current CoM
//     position, so magnitudes stay below a metre whatever the walked distance);
//   * the walk state, the applied first ZMP sample, the LIPM update, the push, the box-violation
//     test and all outputs are FP64, and the applied sample is clamped to the FP64 box, so the
//     strict guarantee (realised ZMP inside the box to ~1e-13) is the same as on the FP64 path.
//
// Organisation (one persistent CTA per SM, 128 columns = 128 tensor-memory lanes per CTA):
//   * D[column, row] = A[column, k] * B[row, k]:   A = right-hand sides (columns x contraction index),
//     B = [M | M G Px] (shared by every column, loaded once by bulk copy, FP16 hi and lo images).
//     Contraction slots 152..154 carry the axis state, so the  G Px x  part of the linear term
//     costs nothing (it rides in the K padding) and only  Q (lo + hi) / 2  is added in the epilogue.
//   * one thread (warp 0, lane 0) issues the 30 MMAs (M = 128, N = 160, K = 16) of an iteration and
//     commits them to an mbarrier; the MMAs of an iteration need every row of the previous epilogue,
//     so contraction and epilogue alternate and no warp is set aside for issuing;
//   * all 16 warps are the epilogue: thread (quadrant q = warp % 4, part = warp / 4) owns column
//     32 q + lane (its tensor-memory lane) and rows 40 part .. 40 part + 39.  It reads v from D,
//     t and lo from tensor memory (hi stays in registers), applies relaxation, box projection, dual
//     update and residual, writes t back and publishes the next right-hand side, split hi / lo,
//     straight into the operand layout the tensor core reads (16-byte stores, conflict free);
//   * a column is a thread, so the stop test needs no shuffles: four partial maxima per column meet
//     in shared memory.  Columns freeze individually the moment THEIR test passes, so results do
//     not depend on which other columns share the CTA.
#include "lipm_internal.cuh"
#include "lipm_tcgen05.cuh"

#include <cuda_fp16.h>

namespace lipm {

namespace {

constexpr int kEpiWarps = 16;
constexpr int kParts = 4;
constexpr int kPartRows = 40;
constexpr int kChunks = 5;                           // 8-row chunks per part
constexpr int kThreadsF16 = 32 * kEpiWarps;          // 512 (128 registers per thread)
constexpr int kLboA = (kF16Cols / 8) * 128;          // 2048
constexpr int kLboB = (kF16Tile / 8) * 128;          // 2560
constexpr uint32_t kTmemD = 0, kTmemT = 160, kTmemLo = 320;   // column offsets (512 allocated)

struct F16Params {
    ColumnBatch cb;
    int N;
    int max_iter;
    float rho, two_rho, alpha, qhalf;
    float eps_stop;          // eps / max(1 + alpha, rho alpha)
    double viol_tol;
    double T, c2, c3, hg, px02, pu00;
    const unsigned char* bimg;   // FP16 hi image then lo image of B
    const double* gpx;           // [NP,3]
    unsigned int* tile_counter;
    unsigned long long* stats;
    unsigned int num_units;
};

struct Control {
    unsigned long long bar_load, bar_mma;
    uint32_t tmem_base;
    int unit, max_ticks, pad;
    float gpx[kF16Tile * 3];
    double X[kF16Cols][3];
    uint32_t Res[kParts][kF16Cols];
    uint32_t Qm[kParts][kF16Cols];
};

constexpr size_t kSmemOperands = 2 * (size_t)kF16OperandBytesB + 2 * (size_t)kF16OperandBytesA;   // 184,320
constexpr size_t kSmemF16 = kSmemOperands + sizeof(Control);

using namespace tc;

// Eight consecutive contraction slots of one column, split into FP16 hi + lo, as two 16-byte stores
// into the operand images (K chunk kc, column s: byte offset kc * LBO + s * 16).
__device__ __forceinline__ void publish_rhs(unsigned char* sAh, unsigned char* sAl, int kc, int s, const float (&r)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(r[2 * e], r[2 * e + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(r[2 * e] - back.x, r[2 * e + 1] - back.y);
        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const int off = kc * kLboA + s * 16;
    *reinterpret_cast<uint4*>(sAh + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sAl + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kThreadsF16, 1) k_strict_rollout_f16(const F16Params P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sBh = smem;
    unsigned char* sBl = smem + kF16OperandBytesB;
    unsigned char* sAh = smem + 2 * kF16OperandBytesB;
    unsigned char* sAl = sAh + kF16OperandBytesA;
    Control& S = *reinterpret_cast<Control*>(smem + kSmemOperands);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, part = warp >> 2;
    const int s = quad * 32 + lane;                  // column = tensor-memory lane
    const int N = P.N;
    const ColumnBatch& B = P.cb;

    // ---- one-time setup: barriers, tensor memory, the shared operand B, G Px -------------------
    if (tid == 0) {
        mbar_init(&S.bar_load, 1);
        mbar_init(&S.bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < kF16Tile * 3; i += kThreadsF16) S.gpx[i] = (i < N * 3) ? (float)P.gpx[i] : 0.f;
    // zero the operand image of the right-hand sides once (rows / slots never written stay zero)
    for (int i = tid; i < 2 * kF16OperandBytesA / 16; i += kThreadsF16)
        reinterpret_cast<uint4*>(sAh)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_expect_tx(&S.bar_load, 2u * kF16OperandBytesB);
        constexpr uint32_t chunk = 2u * kF16OperandBytesB / 4;     // 25,600 B
        for (int c = 0; c < 4; ++c) bulk_load(sBh + c * chunk, P.bimg + c * chunk, chunk, &S.bar_load);
    }
    mbar_wait(&S.bar_load, 0);
    const uint32_t tmem = S.tmem_base + ((uint32_t)(32 * quad) << 16);   // this warp's lane quadrant
    // D = FP32, A = B = FP16, both K-major, N = 160, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kF16Tile >> 3) << 17) | ((uint32_t)(kF16Cols >> 4) << 24);

    // tensor memory starts undefined: clear t and lo once so padding rows read as finite zeros
    {
        uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
            tmem_st8(tmem + kTmemT + part * kPartRows + ch * 8, z);
            tmem_st8(tmem + kTmemLo + part * kPartRows + ch * 8, z);
        }
        tmem_wait_st();
    }

    float hi[kPartRows];          // upper bounds of this thread's rows (tick-relative, FP32)
    uint32_t mma_parity = 0;

    for (;;) {
        // ---- fetch a 128-column unit (y-axis units, which iterate longer, first) ---------------
        __syncthreads();
        if (tid == 0) {
            S.unit = (int)atomicAdd(P.tile_counter, 1u);
            S.max_ticks = 0;
        }
        __syncthreads();
        const unsigned int unit = (unsigned int)S.unit;
        if (unit >= P.num_units) break;
        const int c = (B.unit_order ? B.unit_order[unit] : (int)(P.num_units - 1u - unit)) * kF16Cols + s;

        // column description (every part of a column decodes it; part 0 owns the outputs)
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
            float thr = 0.f;
            {
                // ================= tick prologue =================
                const double x0d = S.X[s][0];
                const float x1f = (float)S.X[s][1], x2f = (float)S.X[s][2];
                const float shift = (float)(origin - x0d);     // re-reference the warm start
                origin = x0d;
                // phase 1: bound window of this tick (relative to the origin); warm start = previous
                // tick's t moved up one row, staged through the (idle) accumulator columns
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    if (part == kParts - 1 && ch == kChunks - 1) continue;      // slots 152..159: axis state
                    const int jb = part * kPartRows + ch * 8;
                    uint32_t lov[8], tv[8], tu[8];
                    if (i > 0) {
                        tmem_ld8(tmem + kTmemT + jb + 1, tv);      // rows jb+1 .. jb+8 of the previous tick
                        tmem_ld8(tmem + kTmemT + jb, tu);          // (row N-1 keeps its own value)
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
                    tmem_st8(tmem + kTmemLo + jb, lov);
                    if (i > 0) tmem_wait_ld();
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
                    tmem_st8(tmem + kTmemD + jb, tv);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncthreads();          // every part has read the old t before anyone overwrites it
                tc_fence_after();
                // phase 2: t <- staged value; |q|_inf; first right-hand side
                uint32_t qm = 0u;
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    const int jb = part * kPartRows + ch * 8;
                    float r[8];
                    if (part == kParts - 1 && ch == kChunks - 1) {
                        // contraction slots 152, 153, 154 = axis state relative to the origin
#pragma unroll
                        for (int e = 0; e < 8; ++e) r[e] = 0.f;
                        r[1] = live ? x1f : 0.f;
                        r[2] = live ? x2f : 0.f;
                    } else {
                        uint32_t tv[8], lov[8];
                        tmem_ld8(tmem + kTmemD + jb, tv);
                        tmem_ld8(tmem + kTmemLo + jb, lov);
                        tmem_wait_ld();
                        tmem_st8(tmem + kTmemT + jb, tv);
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
                    publish_rhs(sAh, sAl, part * kChunks + ch, s, r);
                }
                tmem_wait_st();
                S.Qm[part][s] = qm;
                proxy_fence();
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            bool done = !live;
            bool failed = false;               // stopped without converging (residual stalled far above the threshold)
            uint32_t checkpoint = 0xffffffffu; // residual 32 iterations ago
            int iters = 0;
            {
                uint32_t m = S.Qm[0][s];
#pragma unroll
                for (int p4 = 1; p4 < kParts; ++p4) m = max(m, S.Qm[p4][s]);
                thr = P.eps_stop * fmaxf(1.f, __uint_as_float(m));
            }

            // ================= ADMM =================
            for (int k = 0; k < P.max_iter; ++k) {
                if (warp == 0) {
                    if (lane == 0) {
                        const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);
                        const uint32_t d = S.tmem_base + kTmemD;
#pragma unroll 1
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t a = pass == 1 ? aL : aH, bb = pass == 2 ? bL : bH;
#pragma unroll
                            for (int ks = 0; ks < kF16Tile / 16; ++ks)
                                umma_f16(d, operand_desc(a + ks * 2 * kLboA, kLboA), operand_desc(bb + ks * 2 * kLboB, kLboB),
                                         idesc, (pass | ks) ? 1u : 0u);
                        }
                        umma_commit(&S.bar_mma);
                    }
                    __syncwarp();
                }
                {
                    mbar_wait(&S.bar_mma, mma_parity);
                    tc_fence_after();
                    uint32_t rd = 0u;
                    // tcgen05.ld / st are warp-collective: the warp runs the epilogue while any of its
                    // columns is still iterating; converged columns keep t (and republish the same r)
                    if (__any_sync(0xffffffffu, !done)) {
                        const float relax = done ? 0.f : P.alpha;      // a converged column keeps its t (and republishes the same r)
#pragma unroll
                        for (int ch = 0; ch < kChunks; ++ch) {
                            if (part == kParts - 1 && ch == kChunks - 1) continue;
                            const int jb = part * kPartRows + ch * 8;
                            uint32_t vd[8], vt[8], vl[8];
                            tmem_ld8(tmem + kTmemD + jb, vd);
                            tmem_ld8(tmem + kTmemT + jb, vt);
                            tmem_ld8(tmem + kTmemLo + jb, vl);
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
                            tmem_st8(tmem + kTmemT + jb, vt);
                            publish_rhs(sAh, sAl, part * kChunks + ch, s, r);
                        }
                        tmem_wait_st();
                        rd = done ? 0u : rd;
                    }
                    S.Res[part][s] = rd;
                    proxy_fence();
                }
                mma_parity ^= 1u;
                tc_fence_before();
                __syncthreads();          // new right-hand sides and residual partials are visible; D is consumed
                bool mine = true;
                {
                    if (!done) {
                        uint32_t m = S.Res[0][s];
#pragma unroll
                        for (int p4 = 1; p4 < kParts; ++p4) m = max(m, S.Res[p4][s]);
                        ++iters;
                        done = __uint_as_float(m) <= thr;
                        // A column whose residual has not halved over 32 iterations while still 16x above its
                        // threshold is not going to make it (a diverged walk: FP32 noise on a huge linear term);
                        // give up now instead of dragging the 128-column unit to max_iter.  A healthy solve
                        // contracts by ~0.55 per iteration.
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
                const int stop = __syncthreads_and(mine ? 1 : 0);
                tc_fence_after();
                if (stop) break;
            }

            // ================= apply the first sample (FP64) =================
            if (B.zmp_plan) {
                const bool emit = live && i == nticks - 1;
                double* pz = B.zmp_plan + (long long)(emit ? c : 0) * N;
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    if (part == kParts - 1 && ch == kChunks - 1) continue;
                    const int jb = part * kPartRows + ch * 8;
                    uint32_t vt[8], vl[8];
                    tmem_ld8(tmem + kTmemT + jb, vt);
                    tmem_ld8(tmem + kTmemLo + jb, vl);
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
            if (part == 0) {        // warp-collective loads first, then the columns that are live
                tmem_ld8(tmem + kTmemT, vt0);
                tmem_ld8(tmem + kTmemLo, vl0);
                tmem_wait_ld();
            }
            if (part == 0 && live) {
                const uint32_t (&vt)[8] = vt0;
                const uint32_t (&vl)[8] = vl0;
                const float t0 = __uint_as_float(vt[0]), l0f = __uint_as_float(vl[0]), h0f = hi[0];
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
            tc_fence_before();
            __syncthreads();              // the new state is visible to every part
            tc_fence_after();
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

int launch_strict_rollout_f16(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (!plan->d_bf16 || plan->N > kMaxReducedHorizon) {
        set_error("reduced-precision strict mode needs a plan created with precision F16X3 and horizon <= %d", kMaxReducedHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    F16Params P;
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
    LIPM_CUDA_TRY(cudaFuncSetAttribute(k_strict_rollout_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemF16));
    const unsigned grid = P.num_units < (unsigned)plan->sm_count ? P.num_units : (unsigned)plan->sm_count;
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    P.tile_counter = state.counter();
    P.stats = state.stats();
    k_strict_rollout_f16<<<grid, kThreadsF16, kSmemF16, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(state.publish(plan->d_stats));
    return LIPM_OK;
}

}  // namespace lipm
