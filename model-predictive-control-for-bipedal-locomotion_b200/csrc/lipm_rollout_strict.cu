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
// Organisation (one persistent CTA per SM):
//   * The CTA holds M once, in shared memory, in A-fragment order (N=150: 184,832 B, fetched by
//     bulk async copies).  It is the only large operand and it never moves again.
//   * The CTA is split into NG independent SOLVER GROUPS of 4 warps.  A group owns 8 columns
//     (one DMMA column block) for ALL their ticks: it pulls an 8-column unit from a global
//     counter, walks it to the end, pulls the next.  Groups only share M; they synchronise
//     among their own 128 threads with named barriers, so one group's epilogue / bound loads /
//     state update overlap the other groups' contractions and the FP64 tensor pipe stays fed.
//     Work is scheduled at 8-column granularity (y-axis units, which iterate longer, first).
//   * Inside a group, warp rg owns row blocks rg*NRB .. rg*NRB+NRB-1 of M (row groups are rotated
//     by the group index so the short last row group lands on a different scheduler per group).
//   * registers : t, lo, hi, q of the thread's accumulator elements (C-fragment layout): the
//     epilogue touches shared memory only to publish the next right-hand side R[column][k];
//   * HBM       : bounds are read once per tick, the 3-vector state and a flag byte are written
//     once per tick.  Nothing else leaves the SM during a walk.
//
// Determinism: a column stops updating the moment ITS residual test passes (columns are frozen
// individually; the group keeps iterating for the others), and every accumulation has a fixed
// order, so a scenario's result is bit-identical whatever batch, unit or GPU it lands in.
#include "lipm_internal.cuh"
#include "lipm_device.cuh"

#include <cstdlib>

namespace lipm {

namespace {

struct StrictParams {
    ColumnBatch cb;
    int N, NP, KB, RBLK, KP;
    int strict_flags;       // 1: report active-set bits
    int max_iter;
    double rho, alpha, eps, viol_tol;
    double stop_factor;     // 1 / max(1 + alpha, rho * alpha)
    double Qhalf;           // Q / 2
    double T, c2, c3, hg;   // dt, dt^2/2, dt^3/6, h/g
    double px02;            // Px[0][2] = dt^2/2 - h/g
    double pu00;            // Pu[0][0]
    const double* mfrag;
    const double* gpx;
    unsigned int* tile_counter;
    unsigned long long* stats;   // [0] group iterations executed (8 column slots each), [1] units
    unsigned int num_units;   // 8-column work units
    unsigned int segments;    // each unit's walk is handed out in this many consecutive tick ranges ...
    int seg_ticks;            // ... of this many ticks (load balance: the tail of a launch is one range deep)
    int* seg_done;            // [num_units] ranges completed (release/acquire hand-over between groups)
    double* seg_state;        // [num_units][seg_stride] iterate, axis state and running outcomes of a unit
    size_t seg_stride;        // doubles
    int active_groups;        // debug: solver groups per CTA that take work (LIPM_STRICT_GROUPS)
    int debug_skip;           // debug (LIPM_STRICT_SKIP): 1 skip contraction, 2 skip epilogue math, 4 skip barrier, 8 never converge
    size_t nfrag;           // doubles in mfrag
};

// Shared memory of one solver group, fixed-size part (compile-time offsets from the group base).
// RG = warps of the group (each owns a contiguous range of row blocks).
template <int RG>
struct GroupFixed {
    double X[kGroupCols * 3];                 // current axis state
    double Imp[kGroupCols];                   // dt*F/m of the push
    double Viol[kGroupCols];                  // running max box violation
    double Lean[kGroupCols];                  // running max |CoM - ZMP|
    double* Com[kGroupCols];                  // &com[b][0][axis] / &zmp[b][0][axis] (nullptr: not requested)
    double* Zmp[kGroupCols];
    const double* Lo[kGroupCols];             // bound series of each column
    const double* Hi[kGroupCols];
    double* Hist[kGroupCols];
    double* Zplan[kGroupCols];
    uint8_t* Flag[kGroupCols];
    uint32_t Res[2][RG][kGroupCols];  // high words: residual partials; |q| partials
    int Nser[kGroupCols], Nticks[kGroupCols], Ftick[kGroupCols], Iter[kGroupCols], NotConv[kGroupCols];
    int Unit[2];                              // [0] unit index, [1] ticks of the longest column
};

// Layout: [ mbarrier | NG x GroupFixed | G Px | NG x R[8][KP] | M fragments (resident variants) ]
template <int RG>
__host__ __device__ inline size_t smem_fixed_bytes(int NG) {
    return 16 + (size_t)NG * ((sizeof(GroupFixed<RG>) + 15) & ~(size_t)15);
}

// NRB row blocks per warp, NG solver groups per CTA, NPC compile-time padded horizon (0: run time),
// RG warps per group, STREAM: matrix read from L2 instead of shared memory.
template <int NRB, int NG, int NPC, int RG, bool STREAM>
__global__ void __launch_bounds__(32 * RG * NG, 1) k_strict_rollout(const StrictParams P) {
    constexpr int kGroupThreads = 32 * RG;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    // Geometry: compile-time for the specialised horizons (NPC > 0), run-time otherwise.
    const int N = P.N;
    const int NP = NPC > 0 ? NPC : P.NP;
    const int KB = NP / 4, RBLK = NP / 8, KP = NP + 4;
    const int aStride = KB * 32;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int grp = warp / RG;                    // solver group
    const int gtid = tid & (kGroupThreads - 1);   // thread index inside the group
    const int rg = (warp + grp) & (RG - 1);       // row group (rotated per group)
    const int g = lane >> 2, t = lane & 3;
    const ColumnBatch& B = P.cb;
    const int bar_id = 1 + grp;
#ifdef LIPM_DEBUG_KNOBS
    const int dbg_skip = P.debug_skip;
#else
    constexpr int dbg_skip = 0;      // the measurement knobs are compiled out of the shipped library
#endif

    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(smem_raw);
    constexpr size_t kFixedStride = (sizeof(GroupFixed<RG>) + 15) & ~(size_t)15;
    GroupFixed<RG>& S = *reinterpret_cast<GroupFixed<RG>*>(smem_raw + 16 + grp * kFixedStride);
    double* sGPx = reinterpret_cast<double*>(smem_raw + 16 + NG * kFixedStride);
    double* sR = sGPx + NP * 3 + grp * (kGroupCols * KP);     // this group's right-hand sides
    const double* sM = STREAM ? P.mfrag : sGPx + NP * 3 + NG * (kGroupCols * KP);

    // ---- stationary operands: iteration matrix and G Px, bulk copies issued by one thread -------
    if (tid == 0) mbar_init(sBar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t m_bytes = (uint32_t)((size_t)RBLK * KB * 32 * sizeof(double));
        const uint32_t g_bytes = (uint32_t)(NP * 3 * sizeof(double));
        mbar_expect_tx(sBar, (STREAM ? 0u : m_bytes) + g_bytes);
        if (!STREAM) {
            unsigned char* dstM = reinterpret_cast<unsigned char*>(sGPx + NP * 3 + NG * (kGroupCols * KP));
            const uint32_t chunk = 32768;
            for (uint32_t off = 0; off < m_bytes; off += chunk) {
                const uint32_t n = (m_bytes - off < chunk) ? (m_bytes - off) : chunk;
                bulk_load(dstM + off, reinterpret_cast<const unsigned char*>(P.mfrag) + off, n, sBar);
            }
        }
        bulk_load(sGPx, P.gpx, g_bytes, sBar);
    }
    mbar_wait(sBar, 0);
    // From here on the groups never meet again: only group barriers below.

    // ---- this thread's accumulator elements ----------------------------------------------------
    // element (r, e): row j = (rg*NRB + r)*8 + g, local column lc0 + e = 2t + e
    int nrb = RBLK - rg * NRB;
    nrb = nrb < 0 ? 0 : (nrb > NRB ? NRB : nrb);
    const int row0 = rg * NRB * 8 + g;
    const int lc0 = 2 * t;
    const double* aBase = sM + (size_t)(rg * NRB) * aStride + lane;
    const double* bBase = sR + g * KP + t;
    double* rOut0 = sR + lc0 * KP + row0;          // element (r, e) lives at rOut0[e*KP + r*8]
    const bool is_head = (rg == 0) && (g == 0);    // holds preview row 0 of columns lc0, lc0+1
    const double two_rho = 2.0 * P.rho;

    // ADMM state of this thread's elements: tv = relaxed iterate + dual (z = clip(tv), w = tv - z)
    double tv[NRB][2], lo[NRB][2], hi[NRB][2], q[NRB][2];

    if (grp >= P.active_groups) return;   // measurement knob; all groups by default

    for (;;) {
        // ---- fetch an 8-column unit (y-axis units, which iterate longer, are handed out first) --
        group_sync<kGroupThreads>(bar_id);
        if (gtid == 0) {
            S.Unit[0] = (int)atomicAdd(P.tile_counter, 1u);
            S.Unit[1] = 0;
        }
        group_sync<kGroupThreads>(bar_id);
        // Queue entry = (tick range, unit): all first ranges come first, so the range a unit continues from
        // was handed out num_units entries earlier, to a group that is running.
        const unsigned int entry = (unsigned int)S.Unit[0];
        if (entry >= P.num_units * P.segments) break;
        const unsigned int seg = entry / P.num_units;
        const unsigned int unit = entry - seg * P.num_units;
        const int unit_id = B.unit_order ? B.unit_order[unit] : (int)(P.num_units - 1u - unit);
        const int col0 = unit_id * kGroupCols;

        if (gtid < kGroupCols) {
            const int c = col0 + gtid;
            const bool valid = c < B.ncols;
            const int axis = valid ? c / B.B : 0;
            int b = valid ? c - axis * B.B : 0;
            if (valid && B.order) b = B.order[b];
            int n = valid ? (B.n_steps ? B.n_steps[b] : B.n_uniform) : 0;
            n = n > B.n_max ? B.n_max : n;
            int ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
            if (!valid || n < 1 || ticks < 0) ticks = 0;
            const long long bz = B.bounds_index ? B.bounds_index[b] : ((B.bounds_batch == 1) ? 0 : b);
            const double* plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
            const double* phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
            double* ph = axis ? B.hist[1] : B.hist[0];
            if (ph) ph += (long long)b * B.hist_scn_stride;
            uint8_t* pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (valid) {
                const double* s = (axis ? B.init[1] : B.init[0]) + (long long)b * 3;
                x0 = s[0]; x1 = s[1]; x2 = s[2];
            }
            double viol = 0.0, lean = 0.0;
            double* pcom = (valid && B.com) ? B.com + ((long long)b * B.n_max) * 2 + axis : nullptr;
            double* pzmp = (valid && B.zmp) ? B.zmp + ((long long)b * B.n_max) * 2 + axis : nullptr;
            if (valid && B.write_initial && ticks > 0 && seg == 0) {
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
            const bool push = valid && axis == 1 && B.impulse != nullptr;
            S.Lo[gtid] = plo; S.Hi[gtid] = phi; S.Hist[gtid] = ph; S.Flag[gtid] = pf;
            S.Com[gtid] = pcom; S.Zmp[gtid] = pzmp; S.Lean[gtid] = lean;
            S.Zplan[gtid] = (valid && B.zmp_plan) ? B.zmp_plan + (long long)c * N : nullptr;
            S.Nser[gtid] = n; S.Nticks[gtid] = ticks;
            S.Imp[gtid] = push ? B.impulse[b] : 0.0;
            S.Ftick[gtid] = push ? (B.force_tick ? B.force_tick[b] : -1) : -1;
            S.X[gtid * 3 + 0] = x0; S.X[gtid * 3 + 1] = x1; S.X[gtid * 3 + 2] = x2;
            S.Viol[gtid] = viol; S.Iter[gtid] = 0; S.NotConv[gtid] = 0;
            atomicMax(&S.Unit[1], ticks);
        }
        group_sync<kGroupThreads>(bar_id);
        const int max_ticks = S.Unit[1];
        unsigned int group_iters = 0;
        const int tick_begin = (int)seg * P.seg_ticks;
        const int tick_end = (seg + 1u == P.segments || tick_begin + P.seg_ticks > max_ticks) ? max_ticks : tick_begin + P.seg_ticks;
        if (seg > 0 && tick_begin >= max_ticks) continue;      // the walk ended in an earlier range
        double* carry = P.seg_state + (size_t)unit_id * P.seg_stride;
        if (seg > 0) {
            // continue where the previous range of this unit stopped (elements are filed by row group: the
            // warp that owns a row group differs between the solver groups of a CTA)
            if (gtid == 0) {
                const volatile int* ready = P.seg_done + unit_id;
                while (*ready < (int)seg) __nanosleep(256);
                __threadfence();
            }
            group_sync<kGroupThreads>(bar_id);
#pragma unroll
            for (int r = 0; r < NRB; ++r)
#pragma unroll
                for (int e = 0; e < 2; ++e) tv[r][e] = __ldcg(carry + (r * 2 + e) * kGroupThreads + rg * 32 + lane);
            if (gtid < kGroupCols) {
                const double* c = carry + NRB * 2 * kGroupThreads;
                S.X[gtid * 3 + 0] = __ldcg(c + gtid * 3 + 0);
                S.X[gtid * 3 + 1] = __ldcg(c + gtid * 3 + 1);
                S.X[gtid * 3 + 2] = __ldcg(c + gtid * 3 + 2);
                S.Viol[gtid] = __ldcg(c + 3 * kGroupCols + gtid);
                S.Lean[gtid] = __ldcg(c + 4 * kGroupCols + gtid);
                const int* ci = reinterpret_cast<const int*>(c + 5 * kGroupCols);
                S.Iter[gtid] = __ldcg(ci + gtid);
                S.NotConv[gtid] = __ldcg(ci + kGroupCols + gtid);
                S.Nticks[gtid] = __ldcg(ci + 2 * kGroupCols + gtid);
            }
            group_sync<kGroupThreads>(bar_id);
        }

        for (int i = tick_begin; i < tick_end; ++i) {
            bool live[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) live[e] = i < S.Nticks[lc0 + e];

            // ---- bound window of this tick: row j reads sample min(i + bias + j, n - 1) ------
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double* plo = S.Lo[lc0 + e];
                const double* phi = S.Hi[lc0 + e];
                const int last = S.Nser[lc0 + e] - 1;
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
                group_sync<kGroupThreads>(bar_id);
#pragma unroll
                for (int r = 0; r < NRB; ++r) {
                    const int j = row0 + r * 8;
                    if (r < nrb && j < N) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double shifted = sR[(lc0 + e) * KP + (j + 1 < N ? j + 1 : N - 1)];
                            // last row: keep the primal value, drop the dual
                            tv[r][e] = (j + 1 < N) ? shifted : clipd(shifted, lo[r][e], hi[r][e]);
                        }
                    }
                }
                group_sync<kGroupThreads>(bar_id);
            }

            // ---- linear term, its magnitude (scale of the stop test) and first right-hand side ----
            uint32_t qmag[2] = {0u, 0u};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double x0 = S.X[(lc0 + e) * 3 + 0], x1 = S.X[(lc0 + e) * 3 + 1], x2 = S.X[(lc0 + e) * 3 + 2];
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
                    qmag[e] = max(qmag[e], mag_hi(qv));
                    // rho (z - w) - q with z = clip(t), w = t - z  ->  rho (2z - t) - q
                    const double zc = clipd(tv[r][e], lo[r][e], hi[r][e]);
                    if (r < nrb) rOut0[e * KP + r * 8] = fma(P.rho, 2.0 * zc - tv[r][e], -qv);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) qmag[e] = max(qmag[e], __shfl_xor_sync(0xffffffffu, qmag[e], o));
                if (g == 0) S.Res[1][rg][lc0 + e] = qmag[e];
            }
            group_sync<kGroupThreads>(bar_id);
            // Stop threshold of my two columns (as a high word).  The test uses d = |v - z_old|: the
            // primal residual |v - z_new| <= (1 + alpha) d and the dual one rho |z_new - z_old| <= rho alpha d,
            // so d <= eps * max(1, |q|_inf) / max(1 + alpha, rho alpha) bounds both by eps * max(1, |q|_inf).
            uint32_t thr[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                uint32_t m = S.Res[1][0][lc0 + e];
#pragma unroll
                for (int q4 = 1; q4 < RG; ++q4) m = max(m, S.Res[1][q4][lc0 + e]);
                const double scale = fmax(1.0, __hiloint2double((int)m, 0));
                thr[e] = mag_hi(P.eps * scale * P.stop_factor);
            }

            // ---- ADMM ---------------------------------------------------------------------------
            bool done[2] = {!live[0], !live[1]};
            int iters[2] = {0, 0};
            for (int k = 0; k < P.max_iter; ++k) {
                double acc[NRB][2];
#pragma unroll
                for (int r = 0; r < NRB; ++r) acc[r][0] = acc[r][1] = 0.0;

                long long tc0 = 0;
                if (dbg_skip & 16) tc0 = clock64();
                if (!(dbg_skip & 1)) ContractDispatch<NRB, NRB, NPC / 4, STREAM>::run(nrb, acc, aBase, aStride, bBase, KB);
                if ((dbg_skip & 16) && tid == 0) {
                    // wait for the accumulators, then charge the elapsed cycles to the contraction
                    double sink = 0.0;
#pragma unroll
                    for (int r = 0; r < NRB; ++r) sink += acc[r][0];
                    if (sink == 1.2345e300) S.Viol[0] = sink;
                    atomicAdd(P.stats + 2, (unsigned long long)(clock64() - tc0));
                }

                // relaxation, projection, dual update, residual, next right-hand side
                uint32_t rd[2] = {0u, 0u};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!done[e] && !(dbg_skip & 2)) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r) {
                            if (r >= nrb) continue;
                            const double zo = clipd(tv[r][e], lo[r][e], hi[r][e]);
                            const double d = acc[r][e] - zo;
                            // alpha v + (1 - alpha) zo + (t - zo)  =  t + alpha (v - zo)
                            const double tn = fma(P.alpha, d, tv[r][e]);
                            const double zn = clipd(tn, lo[r][e], hi[r][e]);
                            rd[e] = max(rd[e], mag_hi(d));
                            tv[r][e] = tn;
                            acc[r][e] = fma(two_rho, zn, fma(-P.rho, tn, -q[r][e]));
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#pragma unroll
                    for (int o = 4; o <= 16; o <<= 1) rd[e] = max(rd[e], __shfl_xor_sync(0xffffffffu, rd[e], o));
                    if (g == 0) S.Res[0][rg][lc0 + e] = rd[e];
                }
                if (!(dbg_skip & 4)) group_sync<kGroupThreads>(bar_id);   // every warp of the group has consumed R; partials are visible

                bool all_mine = true;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!done[e]) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r)
                            if (r < nrb) rOut0[e * KP + r * 8] = acc[r][e];
                        uint32_t cd = S.Res[0][0][lc0 + e];
#pragma unroll
                        for (int q4 = 1; q4 < RG; ++q4) cd = max(cd, S.Res[0][q4][lc0 + e]);
                        ++iters[e];
                        done[e] = (cd <= thr[e]) && !(dbg_skip & 8);
                    }
                    all_mine = all_mine && done[e];
                }
                ++group_iters;
                const bool stop = group_sync_and<kGroupThreads>(bar_id, all_mine);   // new R visible; partials consumed
                if ((dbg_skip & 16) && tid == 0) atomicAdd(P.stats + 3, (unsigned long long)(clock64() - tc0));
                if (stop) break;
            }

            // ---- apply the first sample: state update, push, outputs --------------------------
            if (is_head) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int lc = lc0 + e;
                    if (!live[e]) continue;
                    const double x0 = S.X[lc * 3 + 0], x1 = S.X[lc * 3 + 1], x2 = S.X[lc * 3 + 2];
                    const double l0 = lo[0][e], h0 = hi[0][e];
                    const double z0 = clipd(tv[0][e], l0, h0);
                    const double pred = fma(P.px02, x2, fma(P.T, x1, x0));
                    const double jerk = (z0 - pred) / P.pu00;
                    double n0 = fma(P.c3, jerk, fma(P.c2, x2, fma(P.T, x1, x0)));
                    double n1 = fma(P.c2, jerk, fma(P.T, x2, x1));
                    double n2 = fma(P.T, jerk, x2);
                    const double zmp = n0 - P.hg * n2;
                    const double below = l0 - zmp, above = zmp - h0;
                    S.Viol[lc] = fmax(S.Viol[lc], fmax(fmax(below, above), 0.0));
                    S.Lean[lc] = fmax(S.Lean[lc], fabs(n0 - zmp));
                    unsigned f = (below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                 (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u);
                    if (P.strict_flags)
                        f |= (z0 == l0 ? LIPM_FLAG_ACTIVE_LO : 0u) | (z0 == h0 ? LIPM_FLAG_ACTIVE_HI : 0u);
                    if (!done[e]) {
                        // the reference raises at the first QP without a solution (zmp_controller.py:193-194):
                        // the walk of this column ends here
                        f |= LIPM_FLAG_NOT_CONVERGED;
                        S.NotConv[lc] += 1;
                        S.Nticks[lc] = i + 1;
                    }
                    if (i == S.Ftick[lc]) n1 -= S.Imp[lc];
                    const long long sample = i + 1;
                    double* ph = S.Hist[lc];
                    if (ph) {
                        double* dst = ph + (sample - B.hist_index_bias) * 3;
                        dst[0] = n0; dst[1] = n1; dst[2] = n2;
                    }
                    if (S.Flag[lc]) S.Flag[lc][sample * 2] = (uint8_t)f;
                    if (S.Com[lc]) S.Com[lc][sample * 2] = n0;
                    if (S.Zmp[lc]) S.Zmp[lc][sample * 2] = zmp;
                    S.X[lc * 3 + 0] = n0; S.X[lc * 3 + 1] = n1; S.X[lc * 3 + 2] = n2;
                    S.Iter[lc] += iters[e];
                }
            }
            if (B.zmp_plan) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double* pz = S.Zplan[lc0 + e];
                    if (pz && i == S.Nticks[lc0 + e] - 1) {
#pragma unroll
                        for (int r = 0; r < NRB; ++r) {
                            const int j = row0 + r * 8;
                            if (r < nrb && j < N) pz[j] = clipd(tv[r][e], lo[r][e], hi[r][e]);
                        }
                    }
                }
            }
            group_sync<kGroupThreads>(bar_id);
        }

        if (gtid == 0) atomicAdd(P.stats, (unsigned long long)group_iters);
        if (tick_end < max_ticks) {
            // ---- hand the unit over to whoever takes its next range ---------------------------------
#pragma unroll
            for (int r = 0; r < NRB; ++r)
#pragma unroll
                for (int e = 0; e < 2; ++e) __stcg(carry + (r * 2 + e) * kGroupThreads + rg * 32 + lane, tv[r][e]);
            if (gtid < kGroupCols) {
                double* c = carry + NRB * 2 * kGroupThreads;
                __stcg(c + gtid * 3 + 0, S.X[gtid * 3 + 0]);
                __stcg(c + gtid * 3 + 1, S.X[gtid * 3 + 1]);
                __stcg(c + gtid * 3 + 2, S.X[gtid * 3 + 2]);
                __stcg(c + 3 * kGroupCols + gtid, S.Viol[gtid]);
                __stcg(c + 4 * kGroupCols + gtid, S.Lean[gtid]);
                int* ci = reinterpret_cast<int*>(c + 5 * kGroupCols);
                __stcg(ci + gtid, S.Iter[gtid]);
                __stcg(ci + kGroupCols + gtid, S.NotConv[gtid]);
                __stcg(ci + 2 * kGroupCols + gtid, S.Nticks[gtid]);
            }
            __threadfence();
            group_sync<kGroupThreads>(bar_id);
            if (gtid == 0) {
                *reinterpret_cast<volatile int*>(P.seg_done + unit_id) = (int)seg + 1;
            }
            continue;
        }

        // ---- per-column summaries -----------------------------------------------------------------
        if (gtid == 0) atomicAdd(P.stats + 1, 1ull);
        if (gtid < kGroupCols) {
            const int c = col0 + gtid;
            if (c < B.ncols) {
                const int axis = c / B.B;
                const int slot = c - axis * B.B;
                const int b = B.order ? B.order[slot] : slot;
                if (B.max_violation) B.max_violation[b * 2 + axis] = S.Viol[gtid];
                if (B.lean_max) B.lean_max[b * 2 + axis] = S.Lean[gtid];
                if (B.status) B.status[b * 2 + axis] = S.NotConv[gtid];
                if (B.iters) B.iters[b * 2 + axis] = S.Iter[gtid];
            }
        }
    }
}

template <int RG, bool STREAM>
size_t strict_smem_bytes(int NG, int NP) {
    const size_t nfrag = STREAM ? 0 : (size_t)(NP / 8) * (NP / 4) * 32;
    return smem_fixed_bytes<RG>(NG) +
           sizeof(double) * ((size_t)NP * 3 + (size_t)NG * kGroupCols * column_pitch(NP) + nfrag);
}

template <int NRB, int NG, int NPC, int RG = kRowGroups, bool STREAM = false>
int launch_variant(lipm_plan* plan, const StrictParams& P, cudaStream_t stream) {
    const size_t smem = strict_smem_bytes<RG, STREAM>(NG, plan->NP);
    auto kernel = k_strict_rollout<NRB, NG, NPC, RG, STREAM>;
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StrictParams Q = P;
    Q.active_groups = NG;
    Q.debug_skip = 0;
#ifdef LIPM_DEBUG_KNOBS     // measurement builds only (tools/iter_cost.py): never compiled into the shipped library
    if (const char* env = std::getenv("LIPM_STRICT_SKIP")) Q.debug_skip = std::atoi(env) & 31;
    if (const char* env = std::getenv("LIPM_STRICT_GROUPS")) {
        const int v = std::atoi(env);
        if (v >= 1 && v < NG) Q.active_groups = v;
    }
#endif
    Q.num_units = (unsigned)((P.cb.ncols + kGroupCols - 1) / kGroupCols);
    if (Q.num_units == 0) return LIPM_OK;
    const unsigned ctas = (Q.num_units + NG - 1) / NG;
    const unsigned grid = ctas < (unsigned)plan->sm_count ? ctas : (unsigned)plan->sm_count;
    // Tick ranges: with a few units per solver group the launch ends with groups idling for
    // up to a whole walk; cutting the walks into ranges makes that tail one range deep.  The arithmetic of a
    // walk does not depend on where it is cut.
    const int max_ticks = P.cb.ticks_override > 0 ? P.cb.ticks_override : P.cb.n_max - 1;
    const unsigned slots = grid * NG;
    unsigned segments = 1;
    if (Q.num_units > slots) segments = (16u * slots + Q.num_units - 1) / Q.num_units;
#ifdef LIPM_DEBUG_KNOBS
    if (const char* env = std::getenv("LIPM_STRICT_SEGMENTS")) {
        const int v = std::atoi(env);
        if (v >= 1 && v <= 8) segments = (unsigned)v;
    }
#endif
    if (plan->force_segments > 0) segments = (unsigned)plan->force_segments;   // lipm_set_option (tests)
    if (segments > 8u) segments = 8u;
    if (max_ticks < 32 * (int)segments) segments = max_ticks >= 64 ? (unsigned)(max_ticks / 32) : 1u;
    if (segments < 1u) segments = 1u;
    Q.segments = segments;
    Q.seg_ticks = max_ticks > 0 ? (max_ticks + (int)segments - 1) / (int)segments : 1;
    Q.seg_stride = (size_t)NRB * 2 * 32 * RG + 7 * kGroupCols;
    Q.seg_done = nullptr; Q.seg_state = nullptr;
    StreamBuffer carry;          // released on every path out of this function
    if (segments > 1) {
        const size_t done_bytes = ((size_t)Q.num_units * sizeof(int) + 255) & ~(size_t)255;
        const size_t bytes = done_bytes + (size_t)Q.num_units * Q.seg_stride * sizeof(double);
        LIPM_CUDA_TRY(carry.alloc(bytes, stream));
        LIPM_CUDA_TRY(cudaMemsetAsync(carry.ptr, 0, done_bytes, stream));
        Q.seg_done = carry.as<int>();
        Q.seg_state = reinterpret_cast<double*>(carry.as<unsigned char>() + done_bytes);
    }
    LaunchState state;           // work counter + statistics of THIS launch: plans are re-entrant across streams
    LIPM_CUDA_TRY(state.begin(stream));
    Q.tile_counter = state.counter();
    Q.stats = state.stats();
    kernel<<<grid, 32 * RG * NG, smem, stream>>>(Q);
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(state.publish(plan->d_stats));
    return LIPM_OK;
}

}  // namespace

int launch_strict_rollout(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (plan->N > kMaxStrictHorizon) {
        set_error("strict mode: horizon %d exceeds this build's limit %d",
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
    P.stop_factor = 1.0 / std::fmax(1.0 + d.alpha, d.rho * d.alpha);
    P.Qhalf = 0.5 * d.Q;
    P.T = plan->T; P.c2 = std::pow(d.dt, 2) / 2.0; P.c3 = std::pow(d.dt, 3) / 6.0; P.hg = plan->hg;
    P.px02 = P.c2 - P.hg;
    P.pu00 = plan->pu00;
    P.mfrag = plan->d_mfrag; P.gpx = plan->d_gpx;
    P.tile_counter = nullptr;
    P.stats = nullptr;
    P.num_units = 0;
    P.nfrag = (size_t)(plan->NP / 8) * (plan->NP / 4) * 32;

    if (plan->NP == 152) return launch_variant<5, 3, 152>(plan, P, stream);   // N = 145..152 (default 150)
    if (plan->N <= kMaxResidentHorizon) {
        // resident matrix, 4 warps per solver group
        switch ((P.RBLK + kRowGroups - 1) / kRowGroups) {
            case 1: return launch_variant<1, 4, 0>(plan, P, stream);
            case 2: return launch_variant<2, 4, 0>(plan, P, stream);
            case 3: return launch_variant<3, 4, 0>(plan, P, stream);
            case 4: return launch_variant<4, 3, 0>(plan, P, stream);
            case 5: return launch_variant<5, 2, 0>(plan, P, stream);   // N = 153..160 (N <= 152 took the specialised variant): two groups fit beside the matrix
            default: break;
        }
    } else if (P.RBLK <= 8 * 5) {
        // N = 161..320: matrix streamed from L2, 8 warps per group
        switch ((P.RBLK + 7) / 8) {
            case 3: return launch_variant<3, 2, 0, 8, true>(plan, P, stream);
            case 4: return launch_variant<4, 1, 0, 8, true>(plan, P, stream);
            case 5: return launch_variant<5, 1, 0, 8, true>(plan, P, stream);
            default: break;
        }
    } else {
        // N = 321..640: matrix streamed from L2, 16 warps per group
        switch ((P.RBLK + 15) / 16) {
            case 3: return launch_variant<3, 1, 0, 16, true>(plan, P, stream);
            case 4: return launch_variant<4, 1, 0, 16, true>(plan, P, stream);
            case 5: return launch_variant<5, 1, 0, 16, true>(plan, P, stream);
            default: break;
        }
    }
    set_error("strict mode: no kernel variant for %d row blocks", P.RBLK);
    return LIPM_E_UNSUPPORTED;
}

}  // namespace lipm
