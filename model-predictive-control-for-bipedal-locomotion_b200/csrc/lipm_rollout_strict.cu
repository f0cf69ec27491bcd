// Strict (box-constrained) receding-horizon rollout: FP64 ADMM on the FP64 tensor pipe.
//
// Replaces, for a whole batch of scenarios and a whole walk per launch, the strict branch of
// ZMPController.predict_wieber_axis (zmp_controller.py:173-195: cvxpy -> OSQP, rebuilt and cold
// started on every call), the state update (:199), the bound windowing / padding and the push
// of generate_com_trajectory_wieber (:81-106).
//
// Per tick and per column (one axis of one scenario) the QP in ZMP coordinates
//     min 1/2 z'Hz + q'z   s.t.  lo <= z <= hi,     H = Q I + G,  q = -Q (lo+hi)/2 - (G Px) x
// is solved by fixed-rho ADMM with over-relaxation.  With the scaled dual w:
//     v   = M (rho (z - w) - q)            M = (H + rho I)^-1, shared by every column
//     t   = alpha v + (1 - alpha) z + w
//     z+  = clip(t, lo, hi),   w+ = t - z+
// (z, w) is a function of t alone (w is non-zero only where t left the box), so each element
// carries just t between iterations.
// One iteration is therefore one [NP x NP] x [NP x C] FP64 contraction (DMMA.8x8x4) whose
// epilogue does the relaxation, the box projection, the dual update, the residuals and the next
// right-hand side.
//
// Residency (one persistent CTA per SM, C = 8*NCB columns per tile, a tile = all ticks of C walks):
//   shared memory : M in A-fragment order (N=150: 184,832 B, loaded once by bulk async copy),
//                   the right-hand sides R[column][k] (B operand), G Px, per-column bookkeeping;
//   registers     : t, lo, hi, q of the thread's accumulator elements (C-fragment layout), so
//                   the epilogue touches shared memory only to publish the new right-hand side;
//   HBM           : bounds are read once per tick, the 3-vector state and a flag byte are written
//                   once per tick.  Nothing else leaves the SM during a walk.
// Warp (rg, cb) owns row blocks rg*NRB .. rg*NRB+NRB-1 and column block cb.
//
// Determinism: a column stops updating the moment ITS residual test passes (columns are frozen
// individually; the tile keeps iterating for the others), and every accumulation has a fixed
// order, so a scenario's result is bit-identical whatever batch, tile or GPU it lands in.
#include "lipm_internal.cuh"

namespace lipm {

namespace {

struct StrictParams {
    ColumnBatch cb;
    int N, NP, KB, RBLK, KP;
    int strict_flags;       // 1: report active-set bits
    int max_iter;
    double rho, alpha, eps, viol_tol;
    double Qhalf;           // Q / 2
    double T, c2, c3, hg;   // dt, dt^2/2, dt^3/6, h/g
    double px02;            // Px[0][2] = dt^2/2 - h/g
    double pu00;            // Pu[0][0]
    const double* mfrag;
    const double* gpx;
    unsigned int* tile_counter;
    unsigned int num_tiles;
    size_t nfrag;           // doubles in mfrag
};

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
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int NRB, int NCB>
__global__ void __launch_bounds__(32 * kRowGroups * NCB, 1) k_strict_rollout(const StrictParams P) {
    constexpr int C = 8 * NCB;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    // ---- shared memory carve-up -------------------------------------------------------------
    double* sM = reinterpret_cast<double*>(smem_raw);
    double* sGPx = sM + P.nfrag;
    double* sR = sGPx + P.NP * 3;
    double* sX = sR + C * P.KP;
    double* sRes = sX + C * 3;
    double* sImp = sRes + 2 * kRowGroups * C;   // sRes: [rg][C] residuals, then [rg][C] rhs magnitudes
    double* sViol = sImp + C;
    const double** sLo = reinterpret_cast<const double**>(sViol + C);
    const double** sHi = sLo + C;
    double** sHist = reinterpret_cast<double**>(const_cast<double**>(sHi + C));
    double** sZplan = sHist + C;
    uint8_t** sFlag = reinterpret_cast<uint8_t**>(sZplan + C);
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sFlag + C);
    int* sNser = reinterpret_cast<int*>(sBar + 1);
    int* sNticks = sNser + C;
    int* sFtick = sNticks + C;
    int* sIter = sFtick + C;
    int* sNotConv = sIter + C;
    int* sTile = sNotConv + C;   // [0] tile index, [1] ticks of the longest column

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rg = warp / NCB, cb = warp % NCB;
    const int N = P.N, KB = P.KB, KP = P.KP;
    const ColumnBatch& B = P.cb;

    // ---- stationary operands: iteration matrix and G Px, one bulk copy per CTA ----------------
    if (tid == 0) mbar_init(sBar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t m_bytes = (uint32_t)(P.nfrag * sizeof(double));
        const uint32_t g_bytes = (uint32_t)(P.NP * 3 * sizeof(double));
        mbar_expect_tx(sBar, m_bytes + g_bytes);
        const uint32_t chunk = 32768;
        for (uint32_t off = 0; off < m_bytes; off += chunk) {
            const uint32_t n = (m_bytes - off < chunk) ? (m_bytes - off) : chunk;
            bulk_load(reinterpret_cast<unsigned char*>(sM) + off,
                      reinterpret_cast<const unsigned char*>(P.mfrag) + off, n, sBar);
        }
        bulk_load(sGPx, P.gpx, g_bytes, sBar);
    }
    mbar_wait(sBar, 0);

    // ---- this thread's accumulator elements ----------------------------------------------------
    // element (r, e): row j = (rg*NRB + r)*8 + g, local column lc[e] = cb*8 + 2t + e
    int nrb = P.RBLK - rg * NRB;
    nrb = nrb < 0 ? 0 : (nrb > NRB ? NRB : nrb);
    const int row0 = rg * NRB * 8 + g;
    const int lc0 = cb * 8 + 2 * t;
    // A fragments of row block r; row blocks past the end of the matrix alias the last real one
    // (their products are computed and discarded) so that the contraction loop has no branches.
    const double* aTile[NRB];
#pragma unroll
    for (int r = 0; r < NRB; ++r) {
        int I = rg * NRB + r;
        I = I < P.RBLK ? I : P.RBLK - 1;
        aTile[r] = sM + (size_t)I * KB * 32 + lane;
    }
    const double* bBase = sR + (cb * 8 + g) * KP + t;
    double* rOut0 = sR + lc0 * KP + row0;          // element (r, e) lives at rOut0[e*KP + r*8]
    const bool is_head = (rg == 0) && (g == 0);    // holds preview row 0 of columns lc0, lc0+1

    // ADMM state of this thread's elements: tv = relaxed iterate + dual (z = clip(tv), w = tv - z)
    double tv[NRB][2], lo[NRB][2], hi[NRB][2], q[NRB][2];

    for (;;) {
        // ---- fetch a tile (y-axis columns, which iterate longer, are handed out first) -------
        __syncthreads();
        if (tid == 0) {
            sTile[0] = (int)atomicAdd(P.tile_counter, 1u);
            sTile[1] = 0;
        }
        __syncthreads();
        const unsigned int tile = (unsigned int)sTile[0];
        if (tile >= P.num_tiles) break;
        const int col0 = (int)(P.num_tiles - 1u - tile) * C;

        if (tid < C) {
            const int c = col0 + tid;
            const bool valid = c < B.ncols;
            const int axis = valid ? c / B.B : 0;
            const int b = valid ? c - axis * B.B : 0;
            int n = valid ? (B.n_steps ? B.n_steps[b] : B.n_uniform) : 0;
            n = n > B.n_max ? B.n_max : n;
            int ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
            if (!valid || n < 1 || ticks < 0) ticks = 0;
            const long long bz = (B.bounds_batch == 1) ? 0 : b;
            const double* plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
            const double* phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
            double* ph = B.hist[axis] ? B.hist[axis] + (long long)b * B.hist_scn_stride : nullptr;
            uint8_t* pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (valid) {
                const double* s = B.init[axis] + (long long)b * 3;
                x0 = s[0]; x1 = s[1]; x2 = s[2];
            }
            double viol = 0.0;
            if (valid && B.write_initial && ticks > 0) {
                const double zmp = x0 - P.hg * x2;
                const double below = plo[0] - zmp, above = zmp - phi[0];
                viol = fmax(fmax(below, above), 0.0);
                if (ph) { ph[0] = x0; ph[1] = x1; ph[2] = x2; }
                if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                          (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
            }
            const bool push = valid && axis == 1 && B.impulse != nullptr;
            sLo[tid] = plo; sHi[tid] = phi; sHist[tid] = ph; sFlag[tid] = pf;
            sZplan[tid] = (valid && B.zmp_plan) ? B.zmp_plan + (long long)c * N : nullptr;
            sNser[tid] = n; sNticks[tid] = ticks;
            sImp[tid] = push ? B.impulse[b] : 0.0;
            sFtick[tid] = push ? (B.force_tick ? B.force_tick[b] : -1) : -1;
            sX[tid * 3 + 0] = x0; sX[tid * 3 + 1] = x1; sX[tid * 3 + 2] = x2;
            sViol[tid] = viol; sIter[tid] = 0; sNotConv[tid] = 0;
            atomicMax(&sTile[1], ticks);
        }
        __syncthreads();
        const int max_ticks = sTile[1];

        for (int i = 0; i < max_ticks; ++i) {
            bool live[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) live[e] = i < sNticks[lc0 + e];

            // ---- bound window of this tick: row j reads sample min(i + bias + j, n - 1) ------
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* plo = sLo[lc0 + e];
                const double* phi = sHi[lc0 + e];
                const int last = sNser[lc0 + e] - 1;
#pragma unroll
                for (int r = 0; r < NRB; ++r) {
                    const int j = row0 + r * 8;
                    double l = 0.0, h = 0.0;
                    if (live[e] && r < nrb && j < N) {
                        int s = i + B.bound_index_bias + j;
                        s = s > last ? last : s;
                        l = __ldg(plo + (long long)s * B.bound_stride);
                        h = __ldg(phi + (long long)s * B.bound_stride);
                    }
                    lo[r][e] = l; hi[r][e] = h;
                }
            }

            // ---- warm start: previous tick's t shifted by one sample (the dual of the new last
            //      row starts at zero); z and w are re-derived against the new window -------------
            if (i == 0) {
#pragma unroll
                for (int r = 0; r < NRB; ++r)
#pragma unroll
                    for (int e = 0; e < 2; ++e) tv[r][e] = 0.5 * (lo[r][e] + hi[r][e]);
            } else {
#pragma unroll
                for (int r = 0; r < NRB; ++r)
                    if (r < nrb) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) rOut0[e * KP + r * 8] = tv[r][e];
                    }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < NRB; ++r) {
                    const int j = row0 + r * 8;
                    if (r < nrb && j < N) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double shifted = sR[(lc0 + e) * KP + (j + 1 < N ? j + 1 : N - 1)];
                            // last row: keep the primal value, drop the dual
                            tv[r][e] = (j + 1 < N) ? shifted : fmin(fmax(shifted, lo[r][e]), hi[r][e]);
                        }
                    }
                }
                __syncthreads();
            }

            // ---- linear term and first right-hand side ----------------------------------------
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double x0 = sX[(lc0 + e) * 3 + 0], x1 = sX[(lc0 + e) * 3 + 1], x2 = sX[(lc0 + e) * 3 + 2];
#pragma unroll
                for (int r = 0; r < NRB; ++r) {
                    const int j = row0 + r * 8;
                    double qv = 0.0;
                    if (live[e] && r < nrb && j < N) {
                        const double gx = fma(sGPx[j * 3 + 2], x2, fma(sGPx[j * 3 + 1], x1, sGPx[j * 3 + 0] * x0));
                        qv = -fma(P.Qhalf, lo[r][e] + hi[r][e], gx);
                    } else {
                        tv[r][e] = 0.0;
                    }
                    q[r][e] = qv;
                    // rho (z - w) - q with z = clip(t), w = t - z  ->  rho (2z - t) - q
                    const double zc = fmin(fmax(tv[r][e], lo[r][e]), hi[r][e]);
                    if (r < nrb) rOut0[e * KP + r * 8] = fma(P.rho, 2.0 * zc - tv[r][e], -qv);
                }
            }
            __syncthreads();

            // ---- ADMM ---------------------------------------------------------------------------
            bool done[2] = {!live[0], !live[1]};
            int iters[2] = {0, 0};
            for (int k = 0; k < P.max_iter; ++k) {
                double acc[NRB][2];
#pragma unroll
                for (int r = 0; r < NRB; ++r) acc[r][0] = acc[r][1] = 0.0;

                if (__any_sync(0xffffffffu, !done[0] || !done[1])) {
#pragma unroll 2
                    for (int kb = 0; kb < KB; ++kb) {
                        const double bfrag = bBase[kb * 4];
                        double afrag[NRB];
#pragma unroll
                        for (int r = 0; r < NRB; ++r) afrag[r] = aTile[r][kb * 32];
#pragma unroll
                        for (int r = 0; r < NRB; ++r) dmma884(acc[r][0], acc[r][1], afrag[r], bfrag);
                    }
                }

                // residuals and the magnitude of the new right-hand side (stop test is relative to it)
                double res[2] = {0.0, 0.0}, mag[2] = {0.0, 0.0};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!done[e]) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r) {
                            if (r >= nrb) continue;   // aliased row block: product discarded
                            const double v = acc[r][e];
                            const double zo = fmin(fmax(tv[r][e], lo[r][e]), hi[r][e]);
                            const double wo = tv[r][e] - zo;
                            const double tn = fma(P.alpha, v, fma(1.0 - P.alpha, zo, wo));
                            const double zn = fmin(fmax(tn, lo[r][e]), hi[r][e]);
                            res[e] = fmax(res[e], fmax(fabs(v - zn), P.rho * fabs(zn - zo)));
                            tv[r][e] = tn;
                            const double rn = fma(P.rho, 2.0 * zn - tn, -q[r][e]);
                            mag[e] = fmax(mag[e], fabs(rn));
                            acc[r][e] = rn;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#pragma unroll
                    for (int o = 4; o <= 16; o <<= 1) {
                        res[e] = fmax(res[e], __shfl_xor_sync(0xffffffffu, res[e], o));
                        mag[e] = fmax(mag[e], __shfl_xor_sync(0xffffffffu, mag[e], o));
                    }
                }
                if (g == 0) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        sRes[rg * C + lc0 + e] = res[e];
                        sRes[(kRowGroups + rg) * C + lc0 + e] = mag[e];
                    }
                }
                __syncthreads();   // every warp has consumed R; residual partials are visible

                bool all_mine = true;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!done[e]) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r)
                            if (r < nrb) rOut0[e * KP + r * 8] = acc[r][e];
                        double cr = sRes[lc0 + e], cm = sRes[kRowGroups * C + lc0 + e];
#pragma unroll
                        for (int q4 = 1; q4 < kRowGroups; ++q4) {
                            cr = fmax(cr, sRes[q4 * C + lc0 + e]);
                            cm = fmax(cm, sRes[(kRowGroups + q4) * C + lc0 + e]);
                        }
                        ++iters[e];
                        done[e] = (cr <= P.eps * fmax(1.0, cm));
                    }
                    all_mine = all_mine && done[e];
                }
                if (__syncthreads_and(all_mine)) break;   // new R visible; partials consumed
            }

            // ---- apply the first sample: state update, push, outputs --------------------------
            if (is_head) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int lc = lc0 + e;
                    if (!live[e]) continue;
                    const double x0 = sX[lc * 3 + 0], x1 = sX[lc * 3 + 1], x2 = sX[lc * 3 + 2];
                    const double l0 = lo[0][e], h0 = hi[0][e];
                    const double z0 = fmin(fmax(tv[0][e], l0), h0);
                    const double pred = fma(P.px02, x2, fma(P.T, x1, x0));
                    const double jerk = (z0 - pred) / P.pu00;
                    double n0 = fma(P.c3, jerk, fma(P.c2, x2, fma(P.T, x1, x0)));
                    double n1 = fma(P.c2, jerk, fma(P.T, x2, x1));
                    double n2 = fma(P.T, jerk, x2);
                    const double zmp = n0 - P.hg * n2;
                    const double below = l0 - zmp, above = zmp - h0;
                    sViol[lc] = fmax(sViol[lc], fmax(fmax(below, above), 0.0));
                    unsigned f = (below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                 (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u);
                    if (P.strict_flags)
                        f |= (z0 == l0 ? LIPM_FLAG_ACTIVE_LO : 0u) | (z0 == h0 ? LIPM_FLAG_ACTIVE_HI : 0u);
                    if (!done[e]) { f |= LIPM_FLAG_NOT_CONVERGED; sNotConv[lc] += 1; }
                    if (i == sFtick[lc]) n1 -= sImp[lc];
                    const long long sample = i + 1;
                    double* ph = sHist[lc];
                    if (ph) {
                        double* dst = ph + (sample - B.hist_index_bias) * 3;
                        dst[0] = n0; dst[1] = n1; dst[2] = n2;
                    }
                    if (sFlag[lc]) sFlag[lc][sample * 2] = (uint8_t)f;
                    sX[lc * 3 + 0] = n0; sX[lc * 3 + 1] = n1; sX[lc * 3 + 2] = n2;
                    sIter[lc] += iters[e];
                }
            }
            if (B.zmp_plan) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double* pz = sZplan[lc0 + e];
                    if (pz && i == sNticks[lc0 + e] - 1) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r) {
                            const int j = row0 + r * 8;
                            if (r < nrb && j < N) pz[j] = fmin(fmax(tv[r][e], lo[r][e]), hi[r][e]);
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- per-column summaries -----------------------------------------------------------------
        if (tid < C) {
            const int c = col0 + tid;
            if (c < B.ncols) {
                const int axis = c / B.B, b = c - axis * B.B;
                if (B.max_violation) B.max_violation[b * 2 + axis] = sViol[tid];
                if (B.status) B.status[b * 2 + axis] = sNotConv[tid];
                if (B.iters) B.iters[b * 2 + axis] = sIter[tid];
            }
        }
    }
}

template <int NRB, int NCB>
size_t strict_smem_bytes(int NP) {
    constexpr int C = 8 * NCB;
    const size_t nfrag = (size_t)(NP / 8) * (NP / 4) * 32;
    size_t doubles = nfrag + (size_t)NP * 3 + (size_t)C * column_pitch(NP) + C * 3 + 2 * kRowGroups * C + 2 * C;
    size_t bytes = doubles * sizeof(double);
    bytes += 5 * C * sizeof(void*);   // sLo sHi sHist sZplan sFlag
    bytes += sizeof(unsigned long long);
    bytes += (5 * C + 2) * sizeof(int);
    return bytes + 16;
}

template <int NRB, int NCB>
int launch_variant(lipm_plan* plan, const StrictParams& P, cudaStream_t stream) {
    const size_t smem = strict_smem_bytes<NRB, NCB>(plan->NP);
    auto kernel = k_strict_rollout<NRB, NCB>;
    static bool configured = false;   // attribute is per function and per device context
    (void)configured;
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StrictParams Q = P;
    constexpr int C = 8 * NCB;
    Q.num_tiles = (unsigned)((P.cb.ncols + C - 1) / C);
    if (Q.num_tiles == 0) return LIPM_OK;
    const unsigned grid = Q.num_tiles < (unsigned)plan->sm_count ? Q.num_tiles : (unsigned)plan->sm_count;
    LIPM_CUDA_TRY(cudaMemsetAsync(plan->d_tile_counter, 0, sizeof(unsigned int), stream));
    kernel<<<grid, 32 * kRowGroups * NCB, smem, stream>>>(Q);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

}  // namespace

int launch_strict_rollout(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (plan->N > kMaxStrictHorizon) {
        set_error("strict mode: horizon %d exceeds the resident-matrix kernel's limit %d in this build",
                  plan->N, kMaxStrictHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    StrictParams P;
    P.cb = cb;
    P.N = plan->N; P.NP = plan->NP; P.KB = plan->NP / 4; P.RBLK = plan->NP / 8;
    P.KP = column_pitch(plan->NP);
    P.strict_flags = 1;
    P.max_iter = d.max_iter;
    P.rho = d.rho; P.alpha = d.alpha; P.eps = d.eps; P.viol_tol = d.viol_tol;
    P.Qhalf = 0.5 * d.Q;
    P.T = plan->T; P.c2 = std::pow(d.dt, 2) / 2.0; P.c3 = std::pow(d.dt, 3) / 6.0; P.hg = plan->hg;
    P.px02 = P.c2 - P.hg;
    P.pu00 = plan->pu00;
    P.mfrag = plan->d_mfrag; P.gpx = plan->d_gpx;
    P.tile_counter = plan->d_tile_counter;
    P.num_tiles = 0;
    P.nfrag = (size_t)(plan->NP / 8) * (plan->NP / 4) * 32;

    const int nrb = (P.RBLK + kRowGroups - 1) / kRowGroups;
    switch (nrb) {
        case 1: return launch_variant<1, 3>(plan, P, stream);
        case 2: return launch_variant<2, 3>(plan, P, stream);
        case 3: return launch_variant<3, 3>(plan, P, stream);
        case 4: return launch_variant<4, 3>(plan, P, stream);
        case 5: return launch_variant<5, 3>(plan, P, stream);
        default: break;
    }
    set_error("strict mode: no kernel variant for %d row blocks", P.RBLK);
    return LIPM_E_UNSUPPORTED;
}

}  // namespace lipm
