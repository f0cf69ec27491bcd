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
// A push saturates the ZMP: rows 0 .. L-1 are then ALL pinned (a leading block of up to N rows).  Their jerks follow from
// the bounds alone, x_L = Ac^L x + sum_j Ac^(L-1-j) bc b_j (powers of the pinned-stage matrix from a per-plan table: a
// row-parallel dot product), and the rows L .. N-1 form the same QP with horizon N - L started from x_L.  Pu is
// Toeplitz, so the Hessian of that sub-problem is the trailing block H[L:, L:]; the inverses of ALL trailing blocks
// (sum n'^2 = 9 MB at N = 150, built once per plan by bordering) and the matching Kx are per-plan tables, so the
// sub-problem is solved exactly like the full one (its reference term Q Hinv_L zref is a dense product per column and
// tick, since it is no longer shared).  The multipliers of the leading rows are the gradient (H z + q)_a: one product with
// G per sweep.  Only when more than kRowparMaxPins rows are pinned BEHIND the leading block is the column PARKED -- tick,
// state, running outcomes and the set go to a record -- and k_riccati_rollout resumes it from there.
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

struct RowparParams {
    ColumnBatch cb;
    int N;
    int max_sweeps;
    int tick_capacity;          // ticks per run in `u` (n_max - 1)
    double Q, T, c2, c3, hg, cbv, inv_cb, px02, viol_tol, mu_eps;
    double R;
    const double* hl;           // inverses of the trailing blocks H[L:, L:], L = 0 .. N-1, each (N-L)^2 row-major (symmetric)
    const double* kxl;          // Kx of every sub-horizon: [(N-L)][3] per L
    const double* acp;          // [N+1][12]: Ac^m (9, row-major), Ac^m bc (3)
    const double* g;            // [N][N] G = R Pu^-T Pu^-1 (symmetric)
    const double* gpx;          // [NP][3] G Px
    const double* u;            // [runs][tick_capacity][N]  Q H^-1 zref_window
    const double2* pk;          // [runs][n_uniform] (lo, hi)
    const int* n_series;        // [series] samples of each series (nullptr: n_uniform)
    ParkRecord* park;           // [ncols] records of the columns this launch parks (and, for the second stage, resumes)
    int* park_list;             // [ncols] columns parked by this launch
    int* park_count;
    const int* in_list;         // second stage: the columns the first stage parked
    const int* in_count;
    int* work_counter;          // second stage: next item of in_list
    double* s_global;           // second stage: one packed triangle of kMaxRows rows per resident warp
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
template <int KS, int KV>
struct WarpScratch {
    double S[KS * (KS + 1) / 2];            // packed lower triangle: row q at q (q + 1) / 2
    double rr[KV];                          // residual -> y -> lambda
    double dinv[KV];
    double vec[32 * kRowparWords];          // a horizon-long vector every lane reads: zref window, then the plan
    int idx[KV];
};
constexpr int kBlocksSmemPins = 80;         // second stage: rows of the small system (pinned rows behind the leading block, or free
                                            // rows when fewer): with N <= 160 one of the two is always <= 80
constexpr int kMaxRows = 32 * kRowparWords;
constexpr int kFastWarps = 8, kBlocksWarps = 4;   // warps per CTA of the two stages
using FastScratch = WarpScratch<kRowparMaxPins, kRowparMaxPins>;
using BlocksScratch = WarpScratch<kBlocksSmemPins, kMaxRows>;

// offset of the inverse of H[L:, L:] in the table: sum_{l < L} (N - l)^2
__host__ __device__ inline size_t hl_offset(int N, int L) {
    auto sq = [](long long m) { return m * (m + 1) * (2 * m + 1) / 6; };
    return (size_t)(sq(N) - sq(N - L));
}
// offset of the sub-horizon Kx: sum_{l < L} 3 (N - l)
__host__ __device__ inline size_t kxl_offset(int N, int L) { return (size_t)3 * ((size_t)L * N - (size_t)L * (L - 1) / 2); }

// One column's walk.  KR = rows per lane (row r = lane + 32 k).  BLOCKS = false: first stage, every column from tick 0, at
// most kRowparMaxPins pinned rows in all (a short leading block is just a few more pins); lean in registers so that many
// warps are resident.  BLOCKS = true: second stage, resumes a column the first stage parked; leading blocks of any
// length, any number of pins behind them.
template <int KR, bool BLOCKS, class Scratch>
__device__ __forceinline__ void rowpar_column(const RowparParams& P, const int c, Scratch& ws, double* const s_global, const int lane) {
    constexpr int KMAX = BLOCKS ? kMaxRows : kRowparMaxPins;
    const ColumnBatch& B = P.cb;
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

    double kx[KR][3];           // Kx rows of the full horizon (L = 0)
    bool valid[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const int r = lane + 32 * k;
        valid[k] = r < N;
        kx[k][0] = valid[k] ? __ldg(P.kxl + r * 3 + 0) : 0.0;
        kx[k][1] = valid[k] ? __ldg(P.kxl + r * 3 + 1) : 0.0;
        kx[k][2] = valid[k] ? __ldg(P.kxl + r * 3 + 2) : 0.0;
    }

    double viol = 0.0, lean = 0.0;
    int sweeps_total = 0, notconv = 0;
    if (B.write_initial && ticks > 0 && !BLOCKS) {
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
    uint32_t plo = 0u, phi = 0u;          // bit k: my row lane + 32 k is pinned to its lower / upper bound
    bool parked = false;
    int park_tick = 0;
    int i_begin = 0;
    if (BLOCKS) {
        const ParkRecord& rec = P.park[c];
        i_begin = rec.tick; sweeps_total = rec.sweeps; notconv = rec.notconv;
        x0 = rec.x[0]; x1 = rec.x[1]; x2 = rec.x[2]; viol = rec.viol; lean = rec.lean;
#pragma unroll
        for (int k = 0; k < KR; ++k) { plo |= ((rec.m_lo[k] >> lane) & 1u) << k; phi |= ((rec.m_hi[k] >> lane) & 1u) << k; }
    }

    for (int i = i_begin; i < ticks; ++i) {
        if (!compatible) { parked = true; park_tick = i; break; }
        const int base = i + B.bound_index_bias;
        double zu[KR], lo[KR], hi[KR], z[KR];
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            const int r = lane + 32 * k;
            lo[k] = 0.0; hi[k] = 0.0; zu[k] = 0.0; z[k] = 0.0;
            if (valid[k]) {
                const double2 bd = __ldg(pb + min(base + r, last));
                lo[k] = bd.x; hi[k] = bd.y;
            }
        }
        int Lc = -1;                        // leading block the unconstrained plan `zu` was computed for
        uint32_t zmag_hi = 0x3ff00000u;
        bool converged = false;
        for (int sweep = 0; sweep < P.max_sweeps; ++sweep) {
            // ---- the set: leading block, position of my other pinned rows in the list -------------------------------
            uint32_t bal[KR];
            int L = 0, total = 0;
            bool lead = true;
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                bal[k] = __ballot_sync(kFull, ((plo | phi) >> k) & 1u);
                total += __popc(bal[k]);
                if (lead) {
                    const int t = __ffs((int)~bal[k]) - 1;         // -1 when the word is all ones
                    if (t < 0) L += 32; else { L += t; lead = false; }
                }
            }
            if (L > N) L = N;
            if (!BLOCKS) L = 0;                 // first stage: leading rows are ordinary pins
            const int na = total - L;
            if (BLOCKS ? (na > kBlocksSmemPins && N - total > kBlocksSmemPins) : (na > KMAX)) {
                parked = true; park_tick = i;
                if (P.stats && lane == 0) {     // statistics: [2] columns parked (low word: first stage, high: second), [3] sum of L (high) and of pins (low) at the second stage's parks
                    atomicAdd(P.stats + 2, BLOCKS ? (1ull << 32) : 1ull);
                    if (BLOCKS) atomicAdd(P.stats + 3, ((unsigned long long)L << 32) | (unsigned long long)na);
                }
                break;
            }
            ++sweeps_total;
            int myq[KR];
            {
                int seen = 0;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    // pinned rows behind the leading block, in row order
                    const uint32_t behind = (32 * k + 31 < L) ? 0u : (32 * k >= L ? bal[k] : (bal[k] & ~((1u << (L - 32 * k)) - 1u)));
                    myq[k] = seen + __popc(behind & ((1u << lane) - 1u));
                    seen += __popc(behind);
                }
            }
            const double tol = P.mu_eps * P.Q * __hiloint2double((int)zmag_hi, 0);
            bool changed = false;
            uint32_t zm = 0x3ff00000u;
            uint32_t nlo = plo, nhi = phi;
            // the small symmetric positive definite system of a sweep: S (packed lower triangle) x = ws.rr, in place
            auto solve_system = [&](double* S, int ns) {
            // L D L' (right-looking, lane = row; below the diagonsl the entries stay unscaled: M = L d)
            for (int p = 0; p < ns; ++p) {
                const double di = 1.0 / S[p * (p + 1) / 2 + p];
                if (lane == (p & 31)) ws.dinv[p] = di;
                for (int q = lane; q < ns; q += 32) {
                    if (q > p) {
                        double* mine = S + q * (q + 1) / 2;
                        const double lqp = mine[p] * di;
                        int l = p + 1;
                        for (; l + 3 <= q; l += 4) {            // loads first: the stores of a batch cannot alias its loads
                            const double a0 = S[l * (l + 1) / 2 + p], a1 = S[(l + 1) * (l + 2) / 2 + p];
                            const double a2 = S[(l + 2) * (l + 3) / 2 + p], a3 = S[(l + 3) * (l + 4) / 2 + p];
                            const double m0 = mine[l], m1 = mine[l + 1], m2 = mine[l + 2], m3 = mine[l + 3];
                            mine[l] = fma(-lqp, a0, m0); mine[l + 1] = fma(-lqp, a1, m1);
                            mine[l + 2] = fma(-lqp, a2, m2); mine[l + 3] = fma(-lqp, a3, m3);
                        }
                        for (; l <= q; ++l) mine[l] = fma(-lqp, S[l * (l + 1) / 2 + p], mine[l]);
                    }
                }
                __syncwarp();
            }
            // L y = r (column by column), y / d, L' lambda = y
            for (int l = 0; l + 1 < ns; ++l) {
                const double yl = ws.rr[l] * ws.dinv[l];
                for (int q = lane; q < ns; q += 32)
                    if (q > l) ws.rr[q] = fma(-S[q * (q + 1) / 2 + l], yl, ws.rr[q]);
                __syncwarp();
            }
            for (int q = lane; q < ns; q += 32) ws.rr[q] *= ws.dinv[q];
            __syncwarp();
            for (int l = ns - 1; l > 0; --l) {
                const double ll = ws.rr[l];
                for (int q = lane; q < l; q += 32) ws.rr[q] = fma(-S[l * (l + 1) / 2 + q] * ws.dinv[q], ll, ws.rr[q]);
                __syncwarp();
            }
            };
            // Many pinned rows, few free ones (a strong push: runs of saturated rows with short gaps where the box is
            // wider): eliminate the pinned rows instead.  z_A = b_A, H[F, F] z_F = -(q + H[:, A] b_A)_F, and the gradient
            // on the pinned rows is (q + H[:, A] b_A)_A + H[A, F] z_F.  One product with G, a |F| x |F| system.
            const bool free_side = BLOCKS && na > kBlocksSmemPins && N - total <= kBlocksSmemPins;
            if (free_side) {
                int fq[KR], nf = 0;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const uint32_t balf = __ballot_sync(kFull, valid[k] && !(((plo | phi) >> k) & 1u));
                    fq[k] = nf + __popc(balf & ((1u << lane) - 1u));
                    nf += __popc(balf);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KR; ++k)
                    if (valid[k]) ws.vec[lane + 32 * k] = (((plo | phi) >> k) & 1u) ? (((phi >> k) & 1u) ? hi[k] : lo[k]) : 0.0;
                __syncwarp();
                double g0[KR];
#pragma unroll
                for (int k = 0; k < KR; ++k) g0[k] = 0.0;
#pragma unroll 4
                for (int j = 0; j < N; ++j) {
                    const double zj = ws.vec[j];
                    const double* row = P.g + (size_t)j * N + lane;
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (valid[k]) g0[k] = fma(__ldg(row + 32 * k), zj, g0[k]);
                }
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    if (!valid[k]) continue;
                    const int r = lane + 32 * k;
                    const bool pinned = ((plo | phi) >> k) & 1u;
                    const double zp = pinned ? (((phi >> k) & 1u) ? hi[k] : lo[k]) : 0.0;
                    const double gx = fma(__ldg(P.gpx + r * 3), x0, fma(__ldg(P.gpx + r * 3 + 1), x1, __ldg(P.gpx + r * 3 + 2) * x2));
                    g0[k] = fma(P.Q, zp - 0.5 * (lo[k] + hi[k]), g0[k] - gx);
                    if (!pinned) { ws.idx[fq[k]] = r; ws.rr[fq[k]] = -g0[k]; }
                }
                __syncwarp();
                double* const S = ws.S;
                for (int q = lane; q < nf; q += 32) {
                    const double* row = P.g + (size_t)ws.idx[q] * N;
                    for (int l = 0; l <= q; ++l) S[q * (q + 1) / 2 + l] = __ldg(row + ws.idx[l]) + (l == q ? P.Q : 0.0);
                }
                __syncwarp();
                if (nf > 0) solve_system(S, nf);
                for (int q = 0; q < nf; ++q) {
                    const double* row = P.g + (size_t)ws.idx[q] * N + lane;
                    const double zq = ws.rr[q];
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (valid[k] && (((plo | phi) >> k) & 1u)) g0[k] = fma(__ldg(row + 32 * k), zq, g0[k]);
                }
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    if (!valid[k]) continue;
                    const uint32_t bit = 1u << k;
                    if ((plo | phi) & bit) {
                        const bool up = (phi & bit) != 0u;
                        z[k] = up ? hi[k] : lo[k];
                        if (up ? (g0[k] > tol) : (g0[k] < -tol)) { nlo &= ~bit; nhi &= ~bit; changed = true; }
                    } else {
                        z[k] = ws.rr[fq[k]];
                        if (z[k] > hi[k]) { nhi |= bit; changed = true; }
                        else if (z[k] < lo[k]) { nlo |= bit; changed = true; }
                    }
                    zm = max(zm, (uint32_t)__double2hiint(z[k]) & 0x7fffffffu);
                }
                Lc = -1;
            } else {
            const int nsub = N - L;                     // horizon of the sub-problem behind the leading block
            const double* __restrict__ hl = P.hl + hl_offset(N, L);
            // ---- unconstrained plan of the rows behind the leading block --------------------------------------------
            if (L != Lc) {
                Lc = L;
                if (L == 0) {
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (valid[k]) zu[k] = fma(kx[k][0], x0, fma(kx[k][1], x1, fma(kx[k][2], x2, __ldg(pu + (size_t)i * N + lane + 32 * k))));
                } else if (BLOCKS && nsub > 0) {
                    // state behind the block: x_L = Ac^L x + sum_{j < L} Ac^(L-1-j) bc b_j
                    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        const int r = lane + 32 * k;
                        if (r < L) {
                            const double* w = P.acp + (size_t)(L - 1 - r) * 12 + 9;
                            const double bnd = ((phi >> k) & 1u) ? hi[k] : lo[k];
                            t0 = fma(__ldg(w), bnd, t0); t1 = fma(__ldg(w + 1), bnd, t1); t2 = fma(__ldg(w + 2), bnd, t2);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        t0 += __shfl_xor_sync(kFull, t0, o); t1 += __shfl_xor_sync(kFull, t1, o); t2 += __shfl_xor_sync(kFull, t2, o);
                    }
                    const double* ap = P.acp + (size_t)L * 12;
                    const double xl0 = fma(__ldg(ap + 0), x0, fma(__ldg(ap + 1), x1, fma(__ldg(ap + 2), x2, t0)));
                    const double xl1 = fma(__ldg(ap + 3), x0, fma(__ldg(ap + 4), x1, fma(__ldg(ap + 5), x2, t1)));
                    const double xl2 = fma(__ldg(ap + 6), x0, fma(__ldg(ap + 7), x1, fma(__ldg(ap + 8), x2, t2)));
                    // reference term Q Hinv_L zref (not shared any more: dense product) + Kx_L x_L
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (valid[k]) ws.vec[lane + 32 * k] = P.Q * 0.5 * (lo[k] + hi[k]);
                    __syncwarp();
                    const double* kl = P.kxl + kxl_offset(N, L);
                    double acc[KR];
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        const int jr = lane + 32 * k - L;             // my row in the sub-problem
                        acc[k] = (jr >= 0 && jr < nsub) ? fma(__ldg(kl + jr * 3), xl0, fma(__ldg(kl + jr * 3 + 1), xl1, __ldg(kl + jr * 3 + 2) * xl2)) : 0.0;
                    }
#pragma unroll 4
                    for (int cc = 0; cc < nsub; ++cc) {
                        const double zr = ws.vec[L + cc];
                        const double* row = hl + (size_t)cc * nsub - L + lane;
#pragma unroll
                        for (int k = 0; k < KR; ++k) {
                            const int r = lane + 32 * k;
                            if (r >= L && r < N) acc[k] = fma(__ldg(row + 32 * k), zr, acc[k]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < KR; ++k) zu[k] = acc[k];
                }
            }
#pragma unroll
            for (int k = 0; k < KR; ++k) z[k] = zu[k];
            if (na > 0) {
                double* const S = ws.S;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const int r = lane + 32 * k;
                    if ((((plo | phi) >> k) & 1u) && r >= L) {
                        ws.idx[myq[k]] = r - L;
                        ws.rr[myq[k]] = zu[k] - (((phi >> k) & 1u) ? hi[k] : lo[k]);
                    }
                }
                __syncwarp();
                // S = Hinv_L[A, A]
                for (int q = lane; q < na; q += 32) {
                    const double* row = hl + (size_t)ws.idx[q] * nsub;
                    for (int l = 0; l <= q; ++l) S[q * (q + 1) / 2 + l] = __ldg(row + ws.idx[l]);
                }
                __syncwarp();
                solve_system(S, na);
                // correction of my rows
                for (int q = 0; q < na; ++q) {
                    const double* row = hl + (size_t)ws.idx[q] * nsub - L + lane;
                    const double lq = ws.rr[q];
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        const int r = lane + 32 * k;
                        if (r >= L && r < N) z[k] = fma(-lq, __ldg(row + 32 * k), z[k]);
                    }
                }
            }
            // ---- tests: free rows against the box, pinned rows by the sign of their multiplier --------------------------
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                if (!valid[k]) continue;
                const uint32_t bit = 1u << k;
                if ((plo | phi) & bit) {
                    const bool up = (phi & bit) != 0u;
                    z[k] = up ? hi[k] : lo[k];
                    if (lane + 32 * k >= L) {                      // behind the block: the gradient is -lambda
                        const double mu = -ws.rr[myq[k]];
                        if (up ? (mu > tol) : (mu < -tol)) { nlo &= ~bit; nhi &= ~bit; changed = true; }
                    }
                } else {
                    if (z[k] > hi[k]) { nhi |= bit; changed = true; }
                    else if (z[k] < lo[k]) { nlo |= bit; changed = true; }
                }
                zm = max(zm, (uint32_t)__double2hiint(z[k]) & 0x7fffffffu);
            }
            if (BLOCKS && L > 0) {
                // multipliers of the leading rows: mu_a = (H z + q)_a = Q (b_a - zref_a) + (G z)_a - (G Px)_a . x
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KR; ++k)
                    if (valid[k]) ws.vec[lane + 32 * k] = z[k];
                __syncwarp();
                double acc[KR];
#pragma unroll
                for (int k = 0; k < KR; ++k) acc[k] = 0.0;
#pragma unroll 4
                for (int j = 0; j < N; ++j) {
                    const double zj = ws.vec[j];
                    const double* row = P.g + (size_t)j * N + lane;
#pragma unroll
                    for (int k = 0; k < KR; ++k)
                        if (lane + 32 * k < L) acc[k] = fma(__ldg(row + 32 * k), zj, acc[k]);
                }
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const int r = lane + 32 * k;
                    if (r < L) {
                        const uint32_t bit = 1u << k;
                        const bool up = (phi & bit) != 0u;
                        const double gx = fma(__ldg(P.gpx + r * 3), x0, fma(__ldg(P.gpx + r * 3 + 1), x1, __ldg(P.gpx + r * 3 + 2) * x2));
                        const double mu = fma(P.Q, z[k] - 0.5 * (lo[k] + hi[k]), acc[k] - gx);
                        if (up ? (mu > tol) : (mu < -tol)) { nlo &= ~bit; nhi &= ~bit; changed = true; }
                    }
                }
            }
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
            rec.tick = park_tick; rec.sweeps = sweeps_total; rec.notconv = notconv;
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

// first stage: one warp per column
template <int KR>
__global__ void __launch_bounds__(32 * kFastWarps) k_rowpar_rollout(const RowparParams P) {
    __shared__ FastScratch scratch[kFastWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long cl = (long long)blockIdx.x * kFastWarps + wib;
    if (cl >= P.cb.ncols) return;
    rowpar_column<KR, false>(P, (int)cl, scratch[wib], nullptr, lane);
}

// second stage: resident warps pull the parked columns from the list
template <int KR>
__global__ void __launch_bounds__(32 * kBlocksWarps) k_rowpar_blocks(const RowparParams P) {
    extern __shared__ double2 dyn_smem[];
    BlocksScratch* scratch = reinterpret_cast<BlocksScratch*>(dyn_smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* const s_global = P.s_global + ((size_t)blockIdx.x * kBlocksWarps + wib) * (kMaxRows * (kMaxRows + 1) / 2);
    const int count = *P.in_count;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(P.work_counter, 1);
        item = __shfl_sync(kFull, item, 0);
        if (item >= count) break;
        rowpar_column<KR, true>(P, P.in_list[item], scratch[wib], s_global, lane);
        __syncwarp();
    }
}

// Inverses of all trailing blocks of H = Q I + G by bordering, last block first (one CTA, once per plan):
//   inv(H[L:, L:]) = [ 1/s, -v'/s ; -v/s, inv(H[L+1:, L+1:]) + v v'/s ],   v = inv(H[L+1:, L+1:]) g,   s = h - g'v,
// with h = Q + G[L][L], g = G[L+1:, L].
__global__ void __launch_bounds__(1024) k_trailing_inverses(int N, double Q, const double* __restrict__ G, double* __restrict__ hl) {
    extern __shared__ double sm[];       // g (N), v (N), s
    double* gv = sm;
    double* vv = sm + N;
    double* ss = sm + 2 * N;
    // L = N - 1: 1 x 1
    if (threadIdx.x == 0) hl[hl_offset(N, N - 1)] = 1.0 / (Q + G[(size_t)(N - 1) * N + (N - 1)]);
    __syncthreads();
    for (int L = N - 2; L >= 0; --L) {
        const int m = N - L - 1;                              // size of the block already inverted
        const double* old = hl + hl_offset(N, L + 1);
        double* cur = hl + hl_offset(N, L);
        for (int r = threadIdx.x; r < m; r += blockDim.x) gv[r] = G[(size_t)(L + 1 + r) * N + L];
        __syncthreads();
        for (int r = threadIdx.x; r < m; r += blockDim.x) {
            double acc = 0.0;
            for (int cc = 0; cc < m; ++cc) acc = fma(old[(size_t)cc * m + r], gv[cc], acc);     // symmetric: coalesced column read
            vv[r] = acc;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            double acc = 0.0;
            for (int r = threadIdx.x; r < m; r += 32) acc = fma(gv[r], vv[r], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (threadIdx.x == 0) ss[0] = 1.0 / (Q + G[(size_t)L * N + L] - acc);
        }
        __syncthreads();
        const double is = ss[0];
        const int mm = m + 1;
        for (int e = threadIdx.x; e < mm * mm; e += blockDim.x) {
            const int r = e / mm, cc = e - r * mm;
            double val;
            if (r == 0 && cc == 0) val = is;
            else if (r == 0) val = -vv[cc - 1] * is;
            else if (cc == 0) val = -vv[r - 1] * is;
            else val = fma(vv[r - 1] * is, vv[cc - 1], old[(size_t)(r - 1) * m + (cc - 1)]);
            cur[e] = val;
        }
        __syncthreads();
    }
}

// Kx of every sub-horizon: Kx_L = inv(H[L:, L:]) (G[L:, L:] Px[0 : N-L]), one CTA per L
__global__ void __launch_bounds__(256) k_subhorizon_kx(int N, const double* __restrict__ G, const double* __restrict__ px,
                                                       const double* __restrict__ hl, double* __restrict__ kxl) {
    extern __shared__ double sm[];       // (G Px) of the sub-horizon: [m][3]
    const int L = blockIdx.x;
    const int m = N - L;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int d = 0; d < m; ++d) {
            const double gval = G[(size_t)(L + d) * N + (L + r)];
            a0 = fma(gval, px[d * 3 + 0], a0); a1 = fma(gval, px[d * 3 + 1], a1); a2 = fma(gval, px[d * 3 + 2], a2);
        }
        sm[r * 3 + 0] = a0; sm[r * 3 + 1] = a1; sm[r * 3 + 2] = a2;
    }
    __syncthreads();
    const double* inv = hl + hl_offset(N, L);
    double* out = kxl + kxl_offset(N, L);
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int d = 0; d < m; ++d) {
            const double h = inv[(size_t)d * m + r];
            a0 = fma(h, sm[d * 3 + 0], a0); a1 = fma(h, sm[d * 3 + 1], a1); a2 = fma(h, sm[d * 3 + 2], a2);
        }
        out[r * 3 + 0] = a0; out[r * 3 + 1] = a1; out[r * 3 + 2] = a2;
    }
}

}  // namespace

int build_rowpar_tables(lipm_plan* plan) {
    const int N = plan->N;
    if (N > 32 * kRowparWords) return LIPM_OK;         // longer horizons never use the row-parallel kernel
    const lipm_plan_desc_t& d = plan->desc;
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_hl, sizeof(double) * hl_offset(N, N)));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_kx, sizeof(double) * kxl_offset(N, N)));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_acp, sizeof(double) * (size_t)(N + 1) * 12));
    k_trailing_inverses<<<1, 1024, (2 * N + 2) * sizeof(double)>>>(N, d.Q, plan->d_g, plan->d_hl);
    k_subhorizon_kx<<<N, 256, (size_t)N * 3 * sizeof(double)>>>(N, plan->d_g, plan->d_px, plan->d_hl, plan->d_kx);
    LIPM_CUDA_TRY(cudaGetLastError());
    // powers of the pinned-stage matrix Ac = A - B (C A) / (C B) and Ac^m bc, bc = B / (C B)
    const double T = d.dt, c2 = d.dt * d.dt / 2.0, c3 = d.dt * d.dt * d.dt / 6.0, hg = d.h / d.g;
    const double A[3][3] = {{1.0, T, c2}, {0.0, 1.0, T}, {0.0, 0.0, 1.0}};
    const double Bv[3] = {c3, c2, T};
    const double ca[3] = {1.0, T, c2 - hg};            // C A
    const double cb = plan->pu00;                      // C B
    double Ac[3][3], bc[3];
    for (int r = 0; r < 3; ++r) {
        bc[r] = Bv[r] / cb;
        for (int c = 0; c < 3; ++c) Ac[r][c] = A[r][c] - Bv[r] * ca[c] / cb;
    }
    std::vector<double> tab((size_t)(N + 1) * 12);
    double Pm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int m = 0; m <= N; ++m) {
        double* t = &tab[(size_t)m * 12];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) t[r * 3 + c] = Pm[r][c];
            t[9 + r] = Pm[r][0] * bc[0] + Pm[r][1] * bc[1] + Pm[r][2] * bc[2];
        }
        double Nx[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Nx[r][c] = Ac[r][0] * Pm[0][c] + Ac[r][1] * Pm[1][c] + Ac[r][2] * Pm[2][c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Pm[r][c] = Nx[r][c];
    }
    LIPM_CUDA_TRY(cudaMemcpy(plan->d_acp, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    return LIPM_OK;
}

// Can this batch start on the row-parallel kernel?  Walks (not single ticks), horizon within 5 rows per lane, bound
// series shared or indexed (a series per scenario has no shared preview term), preview scratch within bounds.
bool rowpar_eligible(const lipm_plan* plan, const ColumnBatch& cb) {
    if (plan->rowpar_off || !plan->d_hl) return false;
    if (plan->N > 32 * kRowparWords) return false;
    if (cb.ticks_override > 0 || cb.zmp_plan) return false;
    if (!cb.bounds_index && cb.bounds_batch != 1) return false;
    if (cb.bound_axis_offset != 1) return false;
    const size_t u_bytes = (size_t)cb.bounds_batch * 2 * (size_t)(cb.n_max - 1) * plan->N * sizeof(double);
    return u_bytes <= ((size_t)1 << 30);
}

// Runs the batch on the row-parallel kernel; columns whose sets outgrow it are listed in `park` (device) for
// launch_strict_rollout_riccati to resume.
int launch_rowpar(lipm_plan* plan, const ColumnBatch& cb, RowparScratch& rs, cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    const int N = plan->N;
    const int runs = cb.bounds_batch * 2;
    const int tick_capacity = cb.n_max - 1;
    const size_t u_bytes = (size_t)runs * tick_capacity * N * sizeof(double);
    const size_t pk_bytes = (size_t)runs * cb.n_uniform * sizeof(double2);
    const size_t ns_bytes = (((size_t)cb.bounds_batch * sizeof(int)) + 15) & ~(size_t)15;
    const size_t rec_bytes = (size_t)cb.ncols * sizeof(ParkRecord);
    const size_t list_bytes = (((size_t)cb.ncols * 2 + 4) * sizeof(int) + 15) & ~(size_t)15;
    const unsigned blocks2 = (unsigned)std::min<long long>((long long)plan->sm_count * 2, ((long long)cb.ncols + kBlocksWarps - 1) / kBlocksWarps);
    const size_t sg_bytes = (size_t)blocks2 * kBlocksWarps * (kMaxRows * (kMaxRows + 1) / 2) * sizeof(double);
    LIPM_CUDA_TRY(rs.buf.alloc(u_bytes + pk_bytes + ns_bytes + rec_bytes + list_bytes + sg_bytes, stream));
    unsigned char* p = rs.buf.as<unsigned char>();
    double* u = reinterpret_cast<double*>(p); p += u_bytes;
    double2* pk = reinterpret_cast<double2*>(p); p += pk_bytes;
    int* n_series = reinterpret_cast<int*>(p); p += ns_bytes;
    rs.park = reinterpret_cast<ParkRecord*>(p); p += rec_bytes;
    rs.park_count = reinterpret_cast<int*>(p); p += list_bytes;
    rs.park_list = rs.park_count + 4;
    double* s_global = reinterpret_cast<double*>(p);
    LIPM_CUDA_TRY(cudaMemsetAsync(rs.park_count, 0, 4 * sizeof(int), stream));

    {
        const long long total = (long long)runs * cb.n_uniform;
        const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)plan->sm_count * 16);
        k_rowpar_pack<<<blocks ? blocks : 1, 256, 0, stream>>>(cb, pk, runs);
    }
    const int* n_series_arg = nullptr;
    if (cb.n_steps) {
        LIPM_CUDA_TRY(cudaMemsetAsync(n_series, 0xff, (size_t)cb.bounds_batch * sizeof(int), stream));
        k_series_len<<<(cb.B + 255) / 256, 256, 0, stream>>>(cb.B, cb.n_steps, cb.bounds_index, cb.n_max, n_series);
        n_series_arg = n_series;
    }
    {
        const dim3 grid((tick_capacity + kTickTile - 1) / kTickTile, runs);
        k_preview_u<<<grid, 160, (kTickTile + N) * sizeof(double), stream>>>(N, d.Q, cb.n_uniform, cb.bound_index_bias, tick_capacity,
                                                                             n_series_arg, plan->d_hinv, pk, u);
    }
    LIPM_CUDA_TRY(cudaGetLastError());

    RowparParams P;
    P.cb = cb;
    P.N = N;
    P.max_sweeps = plan->riccati_max_sweeps > 0 ? plan->riccati_max_sweeps : 64;
    P.tick_capacity = tick_capacity;
    P.Q = d.Q; P.T = plan->T; P.c2 = d.dt * d.dt / 2.0; P.c3 = d.dt * d.dt * d.dt / 6.0; P.hg = plan->hg;
    P.cbv = plan->pu00; P.inv_cb = 1.0 / plan->pu00; P.px02 = P.c2 - P.hg;
    P.viol_tol = d.viol_tol; P.mu_eps = 1e-12;
    P.R = d.R; P.hl = plan->d_hl; P.kxl = plan->d_kx; P.acp = plan->d_acp; P.g = plan->d_g; P.gpx = plan->d_gpx; P.u = u; P.pk = pk; P.n_series = n_series_arg;
    P.park = rs.park; P.park_list = rs.park_list; P.park_count = rs.park_count;
    P.stats = rs.stats;
    // first stage: all columns; parks into list 0
    P.in_list = nullptr; P.in_count = nullptr; P.work_counter = nullptr; P.s_global = nullptr;
    {
        const unsigned blocks = (unsigned)((cb.ncols + kFastWarps - 1) / kFastWarps);
        k_rowpar_rollout<kRowparWords><<<blocks, 32 * kFastWarps, 0, stream>>>(P);
        LIPM_CUDA_TRY(cudaGetLastError());
    }
    // second stage: the parked columns on resident warps (two CTAs per SM); parks into list 1 what is left for
    // k_riccati_rollout (scenarios whose length differs from their series')
    P.in_list = rs.park_list; P.in_count = rs.park_count;
    P.park_list = rs.park_list + cb.ncols; P.park_count = rs.park_count + 1;
    P.work_counter = rs.park_count + 2;
    P.s_global = s_global;
    {
        const size_t smem = kBlocksWarps * sizeof(BlocksScratch);
        static bool granted = false;
        if (!granted) {
            LIPM_CUDA_TRY(cudaFuncSetAttribute(k_rowpar_blocks<kRowparWords>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            granted = true;
        }
        k_rowpar_blocks<kRowparWords><<<blocks2, 32 * kBlocksWarps, smem, stream>>>(P);
        LIPM_CUDA_TRY(cudaGetLastError());
    }
    rs.park_list = P.park_list; rs.park_count = P.park_count;
    return LIPM_OK;
}

}  // namespace lipm
