// Strict (box-constrained) receding-horizon rollout, FP64, ONE WARP PER COLUMN with the horizon spread over the lanes --
// the first stage of the default FP64 solver for walks whose bound series are shared between scenarios.
//
// Replaces the same reference code as lipm_rollout_riccati.cu (ZMPController.predict_wieber_axis, strict branch,
// zmp_controller.py:173-195; state update :199; windowing / padding / push of generate_com_trajectory_wieber :81-106)
// and runs the same primal-dual active-set iteration, but solves the equality-constrained sub-problems in the
// condensed ZMP coordinates, where every row is independent work:
//   * the unconstrained plan is  z_unc = H^-1 (Q zref + G Px x) = u_i + Kx x  with  u_i = Q H^-1 zref_window(i)  and
//     Kx = H^-1 G Px.  u_i does not depend on the state: it is computed ONCE per bound series, axis and tick
//     (k_preview_u, a small dense product) and shared by every scenario that walks on that series -- the push sweep
//     has one series, the parameter sweep fifteen.  Per column and tick the plan costs three multiply-adds per row;
//   * with a set A of pinned rows (z_A = b_A) the optimum is  z = z_unc - H^-1[:, A] lambda,  H^-1[A, A] lambda =
//     z_unc[A] - b_A:  a |A| x |A| symmetric positive definite system (L D L' in shared memory, one lane per row of L;
//     cond <= cond(H) ~ 8) and |A| multiply-adds per row.  The gradient on a pinned row is -lambda, so the multiplier
//     test needs nothing else.  On the default walk the lateral axis carries ~10 isolated pinned rows (the plan cuts
//     the corner of every support change), the sagittal axis none;
//   * rows that leave the box are pinned, pinned rows whose multiplier has the wrong sign are released, until nothing
//     changes: a fixed point is the KKT point of the reference's strictly convex QP, i.e. its unique minimiser.
// A tick costs a few hundred instructions of a full warp instead of two 150-stage serial sweeps of one thread, and a
// 4096-scenario sweep is 8192 warps: the machine is full.
// A push saturates the ZMP for a long stretch: runs of dozens of pinned rows separated by the short gaps where the box is
// wider (double support).  Dense systems of that size cost more than they save (a second stage with leading-block
// elimination over the trailing-block inverses of H and a free-side formulation was built, agreed to 1e-10 and took
// 140 ms on the push sweep: not shipped), whereas the Riccati recursion takes any set in O(N).  So when a column's set
// outgrows kRowparMaxPins rows it is PARKED -- tick, state, running outcomes and the set go to a record -- and
// k_riccati_rollout resumes it there; when the set has shrunk again (the push is absorbed) that kernel parks the column
// in turn and this kernel finishes the walk.  Launch sequence: all columns here -> parked ones on the Riccati kernel ->
// recovered ones here -> whatever is left on the Riccati kernel.
//
// Determinism: a column's arithmetic does not depend on its neighbours: results are bit-identical whatever batch or GPU
// a scenario lands in.
#include "lipm_internal.cuh"

#include <algorithm>
#include <vector>

namespace lipm {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTickTile = 8;          // ticks per block of k_preview_u
constexpr int kRowparMaxColumns = 16384;

struct RowparParams {
    ColumnBatch cb;
    int N;
    int max_sweeps;
    int tick_capacity;          // ticks per run in `u` (n_max - 1)
    double Q, T, c2, c3, hg, cbv, inv_cb, px02, viol_tol, mu_eps;
    const double* hinv;         // [N][N] inverse Hessian in ZMP coordinates (symmetric)
    const double* kx;           // [N][3] H^-1 G Px
    const double* u;            // [runs][tick_capacity][N]  Q H^-1 zref_window
    const double2* pk;          // [runs][n_uniform] (lo, hi)
    const int* n_series;        // [series] samples of each series (nullptr: n_uniform)
    ParkRecord* park;           // [ncols] records of the columns this launch parks (and, for the second stage, resumes)
    int* park_list;             // [ncols] columns parked by this launch
    int* park_count;
    const int* in_list;         // resume: the columns k_riccati_rollout handed back (nullptr: all columns from tick 0)
    const int* in_count;
    int park_stage;             // stage number written into the records of the columns parked by this launch
    int slot_begin, slot_count; // fresh start: the slots (both axes) of this launch
    unsigned long long* stats;  // [0] sweeps, [1] columns (of the columns that finish here)
};

// (lo, hi) pairs per (series, axis)
__global__ void k_rowpar_pack(const ColumnBatch B, double2* __restrict__ out, int runs) {
    const int ns = B.n_uniform;
    const long long total = (long long)runs * ns;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long run = e / ns;
        const int s = (int)(e - run * ns);
        const long long bz = run >> 1;
        const int axis = (int)(run & 1);
        const long long off = bz * B.bound_scn_stride + (long long)axis * B.bound_axis_offset + (long long)s * B.bound_stride;
        out[e] = make_double2(B.z_min[off], B.z_max[off]);
    }
}

// samples of each series = the walk length of the scenarios that use it (all of them must agree: checked per column)
__global__ void k_series_len(int B, const int32_t* __restrict__ n_steps, const int32_t* __restrict__ bounds_index, int n_max,
                             int* __restrict__ n_series) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int n = n_steps[b];
    n = n > n_max ? n_max : n;
    atomicCAS(n_series + (bounds_index ? bounds_index[b] : 0), -1, n);
}

// u[run][i][r] = Q sum_j Hinv[r][j] zref(min(i + bias + j, n - 1)),  zref = (lo + hi) / 2
__global__ void __launch_bounds__(160) k_preview_u(int N, double Q, int n_uniform, int bias, int tick_capacity,
                                                   const int* __restrict__ n_series, const double* __restrict__ hinv,
                                                   const double2* __restrict__ pk, double* __restrict__ u) {
    extern __shared__ double zs[];      // kTickTile + N window samples
    const int run = blockIdx.y;
    int n = n_series ? n_series[run >> 1] : n_uniform;
    if (n < 1) return;                  // series nobody walks on
    const int i0 = blockIdx.x * kTickTile;
    if (i0 >= n - 1) return;
    const double2* pb = pk + (size_t)run * n_uniform;
    for (int t = threadIdx.x; t < kTickTile + N; t += blockDim.x) {
        const double2 bd = pb[min(i0 + bias + t, n - 1)];
        zs[t] = 0.5 * (bd.x + bd.y);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        double acc[kTickTile];
#pragma unroll
        for (int t = 0; t < kTickTile; ++t) acc[t] = 0.0;
        for (int j = 0; j < N; ++j) {
            const double h = __ldg(hinv + (size_t)j * N + r);          // symmetric: column r read as row-major run
#pragma unroll
            for (int t = 0; t < kTickTile; ++t) acc[t] = fma(h, zs[t + j], acc[t]);
        }
#pragma unroll
        for (int t = 0; t < kTickTile; ++t)
            if (i0 + t < n - 1) u[((size_t)run * tick_capacity + i0 + t) * N + r] = Q * acc[t];
    }
}

// KS = pinned rows whose system fits the shared triangle, KV = pinned rows the vectors can hold
struct WarpScratch {
    double S[kRowparMaxPins * (kRowparMaxPins + 1) / 2];    // packed lower triangle: row q at q (q + 1) / 2
    double rr[kRowparMaxPins];                              // residual -> y -> lambda
    double dinv[kRowparMaxPins];
    int idx[kRowparMaxPins];
};
#ifndef LIPM_ROWPAR_WARPS_PER_BLOCK
#define LIPM_ROWPAR_WARPS_PER_BLOCK 8
#endif
constexpr int kWarpsPerBlock = LIPM_ROWPAR_WARPS_PER_BLOCK;

// KR = rows per lane (row r = lane + 32 k)
template <int KR>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) k_rowpar_rollout(const RowparParams P) {
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    const ColumnBatch& B = P.cb;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const long long cl = (long long)blockIdx.x * kWarpsPerBlock + wib;
    const bool resume = P.in_list != nullptr;
    int c;
    if (resume) {
        if (cl >= *P.in_count) return;
        c = P.in_list[cl];
    } else {
        if (cl >= 2LL * P.slot_count) return;
        const int ax = (int)(cl / P.slot_count);
        c = ax * B.B + P.slot_begin + (int)(cl - (long long)ax * P.slot_count);
    }
    WarpScratch& ws = scratch[wib];
    const int N = P.N;

    const int axis = c / B.B;
    const int slot = c - axis * B.B;
    const int b = B.order ? B.order[slot] : slot;
    int n = B.n_steps ? B.n_steps[b] : B.n_uniform;
    n = n > B.n_max ? B.n_max : n;
    int ticks = n - 1;
    if (n < 1 || ticks < 0) ticks = 0;
    const int series = B.bounds_index ? B.bounds_index[b] : 0;
    const int run = series * 2 + axis;
    // a scenario whose length differs from its series' (ragged walks on a shared series) cannot use the shared preview
    const bool compatible = !P.n_series || P.n_series[series] == n;
    const double2* __restrict__ pb = P.pk + (size_t)run * B.n_uniform;
    const double* __restrict__ pu = P.u + (size_t)run * P.tick_capacity * N;
    double* ph = axis ? B.hist[1] : B.hist[0];
    if (ph) ph += (long long)b * B.hist_scn_stride;
    uint8_t* pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
    double* pcom = B.com ? B.com + ((long long)b * B.n_max) * 2 + axis : nullptr;
    double* pzmp = B.zmp ? B.zmp + ((long long)b * B.n_max) * 2 + axis : nullptr;
    const double* s_init = (axis ? B.init[1] : B.init[0]) + (long long)b * 3;
    double x0 = s_init[0], x1 = s_init[1], x2 = s_init[2];
    const bool push = axis == 1 && B.impulse != nullptr;
    const double impulse = push ? B.impulse[b] : 0.0;
    const int ftick = push ? (B.force_tick ? B.force_tick[b] : -1) : -1;
    const int last = n - 1;

    double kx[KR][3];
    bool valid[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const int r = lane + 32 * k;
        valid[k] = r < N;
        kx[k][0] = valid[k] ? __ldg(P.kx + r * 3 + 0) : 0.0;
        kx[k][1] = valid[k] ? __ldg(P.kx + r * 3 + 1) : 0.0;
        kx[k][2] = valid[k] ? __ldg(P.kx + r * 3 + 2) : 0.0;
    }

    double viol = 0.0, lean = 0.0;
    int sweeps_total = 0, notconv = 0;
    uint32_t plo = 0u, phi = 0u;          // bit k: my row lane + 32 k is pinned to its lower / upper bound
    int i_begin = 0;
    if (resume) {
        // a column k_riccati_rollout hands back: its tick, state, running outcomes and the guess of its set
        const ParkRecord& rec = P.park[c];
        i_begin = rec.tick; sweeps_total = rec.sweeps; notconv = rec.notconv;
        x0 = rec.x[0]; x1 = rec.x[1]; x2 = rec.x[2]; viol = rec.viol; lean = rec.lean;
#pragma unroll
        for (int k = 0; k < KR; ++k) { plo |= ((rec.m_lo[k] >> lane) & 1u) << k; phi |= ((rec.m_hi[k] >> lane) & 1u) << k; }
    } else if (B.write_initial && ticks > 0) {
        const double zmp = x0 - P.hg * x2;
        const double2 b0 = __ldg(pb);
        const double below = b0.x - zmp, above = zmp - b0.y;
        viol = fmax(fmax(below, above), 0.0);
        lean = fabs(x0 - zmp);
        if (lane == 0) {
            if (ph) { ph[0] = x0; ph[1] = x1; ph[2] = x2; }
            if (pcom) pcom[0] = x0;
            if (pzmp) pzmp[0] = zmp;
            if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) | (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
        }
    }

    const double T = P.T, c2 = P.c2, c3 = P.c3, hg = P.hg;
    bool parked = false;
    int park_tick = 0;

    for (int i = i_begin; i < ticks; ++i) {
        if (!compatible) { parked = true; park_tick = i; break; }
        const int base = i + B.bound_index_bias;
        double zu[KR], lo[KR], hi[KR], z[KR];
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const int r = lane + 32 * k;
            if (valid[k]) {
                const double2 bd = __ldg(pb + min(base + r, last));
                lo[k] = bd.x; hi[k] = bd.y;
                zu[k] = fma(kx[k][0], x0, fma(kx[k][1], x1, fma(kx[k][2], x2, __ldg(pu + (size_t)i * N + r))));
            } else {
                lo[k] = 0.0; hi[k] = 0.0; zu[k] = 0.0;
            }
            z[k] = zu[k];
        }
        uint32_t zmag_hi = 0x3ff00000u;
        bool converged = false;
        for (int sweep = 0; sweep < P.max_sweeps; ++sweep) {
            // ---- the set: position of my pinned rows in the list ---------------------------------------------------
            int na = 0;
            int myq[KR];
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const uint32_t bal = __ballot_sync(kFull, ((plo | phi) >> k) & 1u);
                myq[k] = na + __popc(bal & ((1u << lane) - 1u));
                na += __popc(bal);
            }
            if (na > kRowparMaxPins) {
                parked = true; park_tick = i;
                if (P.stats && lane == 0) atomicAdd(P.stats + 2, 1ull);      // statistics: [2] columns parked here
                break;
            }
            ++sweeps_total;
            const double tol = P.mu_eps * P.Q * __hiloint2double((int)zmag_hi, 0);
#pragma unroll
            for (int k = 0; k < KR; ++k) z[k] = zu[k];
            if (na > 0) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    if (((plo | phi) >> k) & 1u) {
                        ws.idx[myq[k]] = lane + 32 * k;
                        ws.rr[myq[k]] = zu[k] - (((phi >> k) & 1u) ? hi[k] : lo[k]);
                    }
                }
                __syncwarp();
                // S = Hinv[A, A]: lane q loads row q
                if (lane < na) {
                    const double* row = P.hinv + (size_t)ws.idx[lane] * N;
                    double* mine = ws.S + lane * (lane + 1) / 2;
                    int l = 0;
                    for (; l + 3 <= lane; l += 4) {               // four loads in flight (L2 / L1 hits)
                        const int i0 = ws.idx[l], i1 = ws.idx[l + 1], i2 = ws.idx[l + 2], i3 = ws.idx[l + 3];
                        const double h0 = __ldg(row + i0), h1 = __ldg(row + i1), h2 = __ldg(row + i2), h3 = __ldg(row + i3);
                        mine[l] = h0; mine[l + 1] = h1; mine[l + 2] = h2; mine[l + 3] = h3;
                    }
                    for (; l <= lane; ++l) mine[l] = __ldg(row + ws.idx[l]);
                }
                __syncwarp();
                // L D L' (right-looking, lane = row; below the diagonal the entries stay unscaled: M = L d)
                for (int p = 0; p < na; ++p) {
                    const double di = 1.0 / ws.S[p * (p + 1) / 2 + p];
                    if (lane == p) ws.dinv[p] = di;
                    if (lane > p && lane < na) {
                        double* mine = ws.S + lane * (lane + 1) / 2;
                        const double lqp = mine[p] * di;
                        int l = p + 1;
                        for (; l + 3 <= lane; l += 4) {           // loads first: the stores of a batch cannot alias its loads
                            const double a0 = ws.S[l * (l + 1) / 2 + p], a1 = ws.S[(l + 1) * (l + 2) / 2 + p];
                            const double a2 = ws.S[(l + 2) * (l + 3) / 2 + p], a3 = ws.S[(l + 3) * (l + 4) / 2 + p];
                            const double m0 = mine[l], m1 = mine[l + 1], m2 = mine[l + 2], m3 = mine[l + 3];
                            mine[l] = fma(-lqp, a0, m0); mine[l + 1] = fma(-lqp, a1, m1);
                            mine[l + 2] = fma(-lqp, a2, m2); mine[l + 3] = fma(-lqp, a3, m3);
                        }
                        for (; l <= lane; ++l) mine[l] = fma(-lqp, ws.S[l * (l + 1) / 2 + p], mine[l]);
                    }
                    __syncwarp();
                }
                // L y = r (column by column), y / d, L' lambda = y
                for (int l = 0; l + 1 < na; ++l) {
                    const double yl = ws.rr[l] * ws.dinv[l];
                    if (lane > l && lane < na) ws.rr[lane] = fma(-ws.S[lane * (lane + 1) / 2 + l], yl, ws.rr[lane]);
                    __syncwarp();
                }
                if (lane < na) ws.rr[lane] *= ws.dinv[lane];
                __syncwarp();
                for (int l = na - 1; l > 0; --l) {
                    const double ll = ws.rr[l];
                    if (lane < l) ws.rr[lane] = fma(-ws.S[l * (l + 1) / 2 + lane] * ws.dinv[lane], ll, ws.rr[lane]);
                    __syncwarp();
                }
                // correction of my rows
                {
                    int q = 0;
                    for (; q + 1 < na; q += 2) {                  // two pinned rows at a time: ten loads in flight
                        const double* ra = P.hinv + (size_t)ws.idx[q] * N + lane;
                        const double* rb = P.hinv + (size_t)ws.idx[q + 1] * N + lane;
                        const double la = ws.rr[q], lb = ws.rr[q + 1];
                        double ha[KR], hb[KR];
#pragma unroll
                        for (int k = 0; k < KR; ++k) { ha[k] = valid[k] ? __ldg(ra + 32 * k) : 0.0; hb[k] = valid[k] ? __ldg(rb + 32 * k) : 0.0; }
#pragma unroll
                        for (int k = 0; k < KR; ++k) z[k] = fma(-lb, hb[k], fma(-la, ha[k], z[k]));
                    }
                    if (q < na) {
                        const double* row = P.hinv + (size_t)ws.idx[q] * N + lane;
                        const double lq = ws.rr[q];
#pragma unroll
                        for (int k = 0; k < KR; ++k)
                            if (valid[k]) z[k] = fma(-lq, __ldg(row + 32 * k), z[k]);
                    }
                }
            }
            // ---- tests: free rows against the box, pinned rows by the sign of their multiplier (-lambda) ----------------
            bool changed = false;
            uint32_t zm = 0x3ff00000u;
            uint32_t nlo = plo, nhi = phi;
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                if (!valid[k]) continue;
                const uint32_t bit = 1u << k;
                if ((plo | phi) & bit) {
                    const bool up = (phi & bit) != 0u;
                    z[k] = up ? hi[k] : lo[k];
                    const double mu = -ws.rr[myq[k]];
                    if (up ? (mu > tol) : (mu < -tol)) { nlo &= ~bit; nhi &= ~bit; changed = true; }
                } else {
                    if (z[k] > hi[k]) { nhi |= bit; changed = true; }
                    else if (z[k] < lo[k]) { nlo |= bit; changed = true; }
                }
                zm = max(zm, (uint32_t)__double2hiint(z[k]) & 0x7fffffffu);
            }
            plo = nlo; phi = nhi;
            zmag_hi = __reduce_max_sync(kFull, zm);
            changed = __any_sync(kFull, changed);
            if (!changed) { converged = true; break; }
        }
        if (parked) break;

        // ---- apply the first sample: state update, push, outputs (every lane carries the state) -----------------------
        {
            const double l0 = __shfl_sync(kFull, lo[0], 0), h0 = __shfl_sync(kFull, hi[0], 0);
            const uint32_t pin0 = __shfl_sync(kFull, (plo & 1u) | ((phi & 1u) << 1), 0);
            const double pred = fma(P.px02, x2, fma(T, x1, x0));      // Px[0] . x
            double z0 = __shfl_sync(kFull, z[0], 0);
            if (pin0) z0 = (pin0 & 2u) ? h0 : l0;
            else z0 = z0 < l0 ? l0 : (z0 > h0 ? h0 : z0);             // a free row is inside the box; keep it there to the last bit
            const double jerk = (z0 - pred) * P.inv_cb;
            double n0 = fma(c3, jerk, fma(c2, x2, fma(T, x1, x0)));
            double n1 = fma(c2, jerk, fma(T, x2, x1));
            double n2 = fma(T, jerk, x2);
            const double zmp = fma(-hg, n2, n0);
            const double below = l0 - zmp, above = zmp - h0;
            viol = fmax(viol, fmax(fmax(below, above), 0.0));
            lean = fmax(lean, fabs(n0 - zmp));
            unsigned f = (below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) | (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u);
            f |= (z0 == l0 ? LIPM_FLAG_ACTIVE_LO : 0u) | (z0 == h0 ? LIPM_FLAG_ACTIVE_HI : 0u);
            if (!converged) { f |= LIPM_FLAG_NOT_CONVERGED; notconv += 1; }
            if (i == ftick) n1 -= impulse;
            const long long sample = i + 1;
            if (lane == 0) {
                if (ph) {
                    double* dst = ph + (sample - B.hist_index_bias) * 3;
                    dst[0] = n0; dst[1] = n1; dst[2] = n2;
                }
                if (pf) pf[sample * 2] = (uint8_t)f;
                if (pcom) pcom[sample * 2] = n0;
                if (pzmp) pzmp[sample * 2] = zmp;
            }
            x0 = n0; x1 = n1; x2 = n2;
            if (!converged) break;
        }
        // ---- guess for the next tick: the window moves up by one sample, so row r takes the status of row r + 1 -----
        {
            const uint32_t dlo = __shfl_down_sync(kFull, plo, 1), dhi = __shfl_down_sync(kFull, phi, 1);
            const uint32_t wlo = __shfl_sync(kFull, plo, 0), whi = __shfl_sync(kFull, phi, 0);
            plo = lane == 31 ? (wlo >> 1) : dlo;
            phi = lane == 31 ? (whi >> 1) : dhi;
#pragma unroll
            for (int k = 0; k < KR; ++k)
                if (!valid[k]) { plo &= ~(1u << k); phi &= ~(1u << k); }
        }
    }

    if (parked) {
        // hand the column to k_riccati_rollout: tick, state, running outcomes, the set as row-bit words
        uint32_t wl[KR], wh[KR];
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            wl[k] = __ballot_sync(kFull, (plo >> k) & 1u);
            wh[k] = __ballot_sync(kFull, (phi >> k) & 1u);
        }
        if (lane == 0) {
            ParkRecord& rec = P.park[c];
            rec.tick = park_tick; rec.sweeps = sweeps_total; rec.notconv = notconv; rec.stage = P.park_stage;
            rec.x[0] = x0; rec.x[1] = x1; rec.x[2] = x2; rec.viol = viol; rec.lean = lean;
#pragma unroll
            for (int k = 0; k < kRowparWords; ++k) { rec.m_lo[k] = 0u; rec.m_hi[k] = 0u; }
#pragma unroll
            for (int k = 0; k < KR; ++k) { rec.m_lo[k] = wl[k]; rec.m_hi[k] = wh[k]; }
            P.park_list[atomicAdd(P.park_count, 1)] = c;
        }
        return;
    }
    if (lane == 0) {
        if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
        if (B.lean_max) B.lean_max[b * 2 + axis] = lean;
        if (B.status) B.status[b * 2 + axis] = notconv;
        if (B.iters) B.iters[b * 2 + axis] = sweeps_total;
        if (P.stats) {
            atomicAdd(P.stats, (unsigned long long)sweeps_total);
            atomicAdd(P.stats + 1, 1ull);
        }
    }
}

// Kx = H^-1 (G Px), once per plan
__global__ void k_rowpar_kx(int N, const double* __restrict__ hinv, const double* __restrict__ gpx, double* __restrict__ kx) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = 0; j < N; ++j) {
        const double h = hinv[(size_t)j * N + r];
        a0 = fma(h, gpx[j * 3 + 0], a0); a1 = fma(h, gpx[j * 3 + 1], a1); a2 = fma(h, gpx[j * 3 + 2], a2);
    }
    kx[r * 3 + 0] = a0; kx[r * 3 + 1] = a1; kx[r * 3 + 2] = a2;
}

}  // namespace

int build_rowpar_tables(lipm_plan* plan) {
    const int N = plan->N;
    if (N > 32 * kRowparWords) return LIPM_OK;         // longer horizons never use the row-parallel kernel
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_kx, sizeof(double) * N * 3));
    k_rowpar_kx<<<(N + 127) / 128, 128>>>(N, plan->d_hinv, plan->d_gpx, plan->d_kx);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

// Can this batch start on the row-parallel kernel?  Walks (not single ticks), horizon within 5 rows per lane, bound
// series shared or indexed (a series per scenario has no shared preview term), preview scratch within bounds.
bool rowpar_eligible(const lipm_plan* plan, const ColumnBatch& cb) {
    if (plan->rowpar_off || !plan->d_kx) return false;
    // One warp per column pays while the batch is small enough to be latency bound on the thread-per-column kernel
    // (8192 columns: 46 against 49 ms with pushes, 9 against 22 ms without); a 32768-column bucket of the parameter
    // sweep fills the machine with full warps of the Riccati kernel and is 25 % faster there.
    if (plan->rowpar_off != -1 && cb.ncols > kRowparMaxColumns) return false;
    if (plan->N > 32 * kRowparWords) return false;
    if (cb.ticks_override > 0 || cb.zmp_plan) return false;
    if (!cb.bounds_index && cb.bounds_batch != 1) return false;
    if (cb.bound_axis_offset != 1) return false;
    const size_t u_bytes = (size_t)cb.bounds_batch * 2 * (size_t)(cb.n_max - 1) * plan->N * sizeof(double);
    return u_bytes <= ((size_t)1 << 30);
}

// Prepares the shared preview terms and the hand-over lists of a batch (rs), once per rollout.
int rowpar_prepare(lipm_plan* plan, const ColumnBatch& cb, RowparScratch& rs, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    const int N = plan->N;
    const int runs = cb.bounds_batch * 2;
    const int tick_capacity = cb.n_max - 1;
    const size_t u_bytes = (size_t)runs * tick_capacity * N * sizeof(double);
    const size_t pk_bytes = (size_t)runs * cb.n_uniform * sizeof(double2);
    const size_t ns_bytes = (((size_t)cb.bounds_batch * sizeof(int)) + 15) & ~(size_t)15;
    const size_t rec_bytes = (size_t)cb.ncols * sizeof(ParkRecord);
    const size_t list_bytes = (((size_t)cb.ncols * kRowparLists * kRowparMaxChunks + 4 * kRowparMaxChunks) * sizeof(int) + 15) & ~(size_t)15;
    LIPM_CUDA_TRY(rs.buf.alloc(u_bytes + pk_bytes + ns_bytes + rec_bytes + list_bytes, stream));
    unsigned char* p = rs.buf.as<unsigned char>();
    rs.u = reinterpret_cast<double*>(p); p += u_bytes;
    rs.pk = reinterpret_cast<double2*>(p); p += pk_bytes;
    int* n_series = reinterpret_cast<int*>(p); p += ns_bytes;
    rs.park = reinterpret_cast<ParkRecord*>(p); p += rec_bytes;
    rs.counts = reinterpret_cast<int*>(p);
    rs.lists = rs.counts + 4 * kRowparMaxChunks;
    rs.ncols = cb.ncols;
    LIPM_CUDA_TRY(cudaMemsetAsync(rs.counts, 0, 4 * kRowparMaxChunks * sizeof(int), stream));
    LIPM_CUDA_TRY(cudaMemsetAsync(rs.park, 0, rec_bytes, stream));
    {
        const long long total = (long long)runs * cb.n_uniform;
        const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)plan->sm_count * 16);
        k_rowpar_pack<<<blocks ? blocks : 1, 256, 0, stream>>>(cb, rs.pk, runs);
    }
    rs.n_series = nullptr;
    if (cb.n_steps) {
        LIPM_CUDA_TRY(cudaMemsetAsync(n_series, 0xff, (size_t)cb.bounds_batch * sizeof(int), stream));
        k_series_len<<<(cb.B + 255) / 256, 256, 0, stream>>>(cb.B, cb.n_steps, cb.bounds_index, cb.n_max, n_series);
        rs.n_series = n_series;
    }
    {
        const dim3 grid((tick_capacity + kTickTile - 1) / kTickTile, runs);
        k_preview_u<<<grid, 160, (kTickTile + N) * sizeof(double), stream>>>(N, d.Q, cb.n_uniform, cb.bound_index_bias, tick_capacity,
                                                                             rs.n_series, plan->d_hinv, rs.pk, rs.u);
    }
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

// One pass of the row-parallel kernel over slot chunk `chunk` of `chunks`: its columns from tick 0 (in_list < 0) or the
// columns of the chunk's hand-over list `in_list`; the columns it parks go to the chunk's list `out_list`.
int launch_rowpar(lipm_plan* plan, const ColumnBatch& cb, const RowparScratch& rs, int chunk, int chunks, int in_list, int out_list,
                  cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    RowparParams P;
    P.cb = cb;
    P.N = plan->N;
    P.max_sweeps = plan->riccati_max_sweeps > 0 ? plan->riccati_max_sweeps : 64;
    P.tick_capacity = cb.n_max - 1;
    P.Q = d.Q; P.T = plan->T; P.c2 = d.dt * d.dt / 2.0; P.c3 = d.dt * d.dt * d.dt / 6.0; P.hg = plan->hg;
    P.cbv = plan->pu00; P.inv_cb = 1.0 / plan->pu00; P.px02 = P.c2 - P.hg;
    P.viol_tol = d.viol_tol; P.mu_eps = 1e-12;
    P.hinv = plan->d_hinv; P.kx = plan->d_kx; P.u = rs.u; P.pk = rs.pk; P.n_series = rs.n_series;
    P.park = rs.park; P.park_list = rs.list(chunk, out_list); P.park_count = rs.count(chunk, out_list);
    P.in_list = in_list < 0 ? nullptr : rs.list(chunk, in_list);
    P.in_count = in_list < 0 ? nullptr : rs.count(chunk, in_list);
    P.park_stage = RowparScratch::stage(chunk, out_list);
    P.slot_begin = (int)((long long)cb.B * chunk / chunks);
    P.slot_count = (int)((long long)cb.B * (chunk + 1) / chunks) - P.slot_begin;
    P.stats = rs.stats;
    const long long warps = 2LL * P.slot_count;          // (a hand-over list holds at most the chunk's columns)
    const unsigned blocks = (unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (blocks == 0) return LIPM_OK;
    k_rowpar_rollout<kRowparWords><<<blocks, 32 * kWarpsPerBlock, 0, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

}  // namespace lipm
