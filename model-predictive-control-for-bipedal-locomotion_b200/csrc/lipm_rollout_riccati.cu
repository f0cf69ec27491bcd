// Strict (box-constrained) receding-horizon rollout, FP64, by an active-set iteration whose sub-problems
// are solved EXACTLY in O(N) by a Riccati sweep -- the default solver of FP64 plans.
//
// Replaces the same reference code as lipm_rollout_strict.cu (ZMPController.predict_wieber_axis, strict branch,
// zmp_controller.py:173-195; the state update :199; the windowing / padding / push of
// generate_com_trajectory_wieber :81-106) and produces the same outputs, but does not touch the condensed
// N x N matrices at all.  The strict QP is an LQ tracking problem of the 3-state model (:15-21):
//     stage i applies jerk J_i:  x_{i+1} = A x_i + B J_i,   previews  z_i = C x_{i+1}   (row i of Px x + Pu J),
//     costs  Q/2 (z_i - zref_i)^2 + R/2 J_i^2   and must keep   zmin_i <= z_i <= zmax_i.
// For a guessed active set (rows pinned to one of their bounds) the problem is an equality-constrained LQ
// problem: a pinned row fixes its own jerk, J_i = (b_i - C A x_i) / (C B), a free row is an ordinary LQ stage.
// One backward sweep gives the cost-to-go 1/2 x'P_i x + p_i'x of every stage (and the feedback gains), one
// forward sweep from the current state gives the plan z and, on pinned rows, the multiplier
// mu_i = d cost / d b_i  (= the gradient (H z + q)_i of the condensed QP).  Then the set is corrected -- free
// rows that left the box are pinned, pinned rows whose multiplier has the wrong sign are released -- and the
// sweeps are repeated until nothing changes (primal-dual active-set iteration).  A fixed point satisfies the KKT
// conditions of the reference's QP, whose minimiser is unique (strictly convex), so the result is THE optimum
// to rounding (1e-14 against the oracle's BVLS), not an eps-approximation.
// Receding horizon makes the guess nearly free: the previous tick's set, moved up by one sample, is right on
// ~97 % of the x-axis ticks and ~60 % of the pushed y-axis ticks of the BASELINE sweep (1.33 sweep pairs per solve
// against 26.8 ADMM iterations of 2 N^2 flops each: ~150x fewer flops per solve at N = 150).
//
// Organisation: ONE THREAD PER COLUMN (one axis of one scenario) for the whole walk; the recursion is a
// dependent chain, so the machine is filled by columns, not by rows.  Registers: axis state, P (6), p (3).
// Per-stage feedback gains (4 doubles per stage) live in the thread's local memory (L1-resident at the BASELINE
// batch), the active set is two bit masks.  Warps carry `lanes` columns each (a power of two <= 32): at small
// batches narrower warps spread the columns over more schedulers and diverge less.
// HBM: bounds read through L1 (the window slides by one sample per tick), 24 B of state + flag / com / zmp per
// solve written once.
//
// Free tail from per-plan tables: the quadratic part P_j of the cost-to-go does not depend on the bounds, only on which
// of the rows j .. N-1 are pinned.  Above the highest pinned row (on most ticks: everywhere, or everywhere past a leading
// pinned block) it is therefore the same for every column and every tick: P_j, the feedback row K_j, the closed-loop
// matrix Acl_j = A - B K_j and B / sigma_j are computed once per plan (build_riccati_tables) and staged in shared memory.
// There a backward stage is only the linear part  v_j = Acl_j'(v_{j+1} - Q zref_j C')  plus the feed-forward
// -kc_j = -(B / sigma_j) . w  (16 multiply-adds, chain depth 3), a forward stage is  s_{j+1} = Acl_j s_j - B kc_j
// (12 multiply-adds, chain depth 3) and the box test; the general recursion (~100 multiply-adds per stage pair) runs
// only from the highest pinned row down to the leading block.
//
// Determinism: a column's arithmetic does not depend on its neighbours, so results are bit-identical whatever
// batch, warp shape or GPU a scenario lands in.
#include "lipm_internal.cuh"

#include <algorithm>
#include <cmath>
#include <memory>
#include <mutex>
#include <vector>

namespace lipm {

namespace {

struct RiccatiParams {
    ColumnBatch cb;
    int N;
    int max_sweeps;
    int lanes;              // columns per warp
    int col_begin, col_end; // columns of this launch (large batches are cut so that the scratch stays bounded)
    double Q, R, T, c2, c3, hg;
    double cbv, inv_cb;     // C B = Pu[0][0] and its inverse
    double px02;            // Px[0][2] = dt^2/2 - h/g
    double r_cb;            // R / (C B)
    double q_half;          // Q / 2
    double q_hg, q_hg2;     // Q h/g, Q (h/g)^2
    double s_const;         // R + Q (C B)^2
    double bb[6];           // B_i B_j (00 01 02 11 12 22, off-diagonals doubled)
    double kv[3];           // pinned row: J = kv . x + b / (C B)
    double bc[3];           // B / (C B)
    double a1[3], a2[3];    // rows 1 and 2 of Ac = A - B (C A) / (C B); row 0 = (h/g) row 2 because C Ac = 0
    double hg2x, hgsq;      // 2 h/g, (h/g)^2
    double rkk[6];          // R kv_i kv_j (00 01 02 11 12 22)
    double viol_tol, mu_eps;
    double2* gains;         // [warp][stage][lane] x 2: per-stage feedback data of the backward sweep (general recursion)
    double* nkc;            // [warp][stage][lane]: feed-forward -kc of the table stages (nullptr: kept in shared memory)
    const double* tab;      // [N + 4][16] per-plan stage tables (kTab* below)
    const double* aux;      // [N + 1][12]: K_j (3), pad, P_j (00 01 02 11 12 22), pad 2
    const double2* pk;      // bound series repacked as (lo, hi) pairs (k_pack_bounds)
    const double* hinv;     // [N][N] inverse Hessian in ZMP coordinates (plan->d_hinv, symmetric): few-pin solves
    int schur_max;          // pinned rows up to which a sweep uses the table solve + dense correction (0: never)
    ParkRecord* park;       // resume mode: records of the columns parked by the row-parallel kernel (nullptr: whole walks)
    int park_stage;         // ... those whose record carries this stage number (columns stay in processing order)
    int* back_list;         // resume mode: columns handed back to the row-parallel kernel once their set is small again
    int* back_count;        //   (nullptr: every column is finished here)
    int back_pins;          // ... small = at most this many pinned rows on kBackStreak consecutive ticks
    int back_stage;         // stage number of the launch that takes them
    int pk_by_column;       // 1: one run of pairs per column of the launch; 0: one per (series, axis)
    unsigned long long* stats;   // [0] sweep pairs, [1] columns, [2] stages of the general recursion, [3] table stages
};

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 2^k as a double from a biased exponent field (no range checks: callers keep it inside [1, 2046])
__device__ __forceinline__ double pow2_from_biased(int e) { return __hiloint2double(e << 20, 0); }
__device__ __forceinline__ int biased_exponent(double x) { return (__double2hiint(x) >> 20) & 0x7ff; }

constexpr int kGainPrefetch = 16;  // stages of look-ahead of the forward sweep's L1 prefetch (a forward stage is ~20 instructions, an L2 hit ~250 cycles)
constexpr int kMaskWords = kMaxStrictHorizon / 32;
#ifndef LIPM_BACK_STREAK
#define LIPM_BACK_STREAK 3
#endif
#ifndef LIPM_BACK_MARGIN
#define LIPM_BACK_MARGIN 4        // a column is handed back when its set is this much below the row-parallel kernel's limit ...
#endif
constexpr int kBackStreak = LIPM_BACK_STREAK;      // ... on this many consecutive ticks
constexpr int kSchurMax = 16;                   // pinned rows a few-pin sweep can take (triangular system in local memory)
constexpr size_t kMaxAutoSmem = 72 * 1024;       // per block: three blocks of two warps still fit one SM
constexpr size_t kMaxDynamicSmem = 200 * 1024;

// Cost-to-go of the backward sweep, kept projectively: 1/2 x'(P/pi)x + (v/pi)'x.  A free stage multiplies the scale by
// its own sigma = pi (R + B'W B) instead of dividing by it, so the reciprocal (needed only for the stored feedback gains)
// is off the dependency chain of the next stage; powers of two keep pi near 1 (exact).
struct CostToGo {
    double P00, P01, P02, P11, P12, P22, v0, v1, v2, pi, ipi;
};

// One FREE stage (ordinary LQ stage with reference zref = (lo + hi) / 2): updates the cost-to-go and returns the
// negated feedback (J = g . (s, 1)) of this stage.
__device__ __forceinline__ void backward_free(const RiccatiParams& P, CostToGo& c, double lo, double hi, double2& g01, double2& g23) {
    const double T = P.T, c2 = P.c2, c3 = P.c3, hg = P.hg;
    const double tq = c.pi * (P.q_half * (lo + hi));     // pi Q zref
    const double w0 = c.v0 - tq, w1 = c.v1, w2 = fma(tq, hg, c.v2);
    const double W00 = fma(c.pi, P.Q, c.P00), W02 = fma(c.pi, -P.q_hg, c.P02), W22 = fma(c.pi, P.q_hg2, c.P22);
    // sigma = pi (R + B'W B) as a quadratic form of P (short chain); beta = W B
    const double sg = fma(c.P00, P.bb[0], fma(c.P01, P.bb[1], fma(c.P02, P.bb[2], c.pi * P.s_const))) +
                      fma(c.P11, P.bb[3], fma(c.P12, P.bb[4], c.P22 * P.bb[5]));
    const double b0 = fma(W02, T, fma(c.P01, c2, W00 * c3));
    const double b1 = fma(c.P12, T, fma(c.P11, c2, c.P01 * c3));
    const double b2 = fma(W22, T, fma(c.P12, c2, W02 * c3));
    const double bw = fma(T, w2, fma(c2, w1, c3 * w0));
    const double u0 = b0, u1 = fma(T, b0, b1), u2 = fma(c2, b0, fma(T, b1, b2));
    // X = A'W A
    const double Y01 = fma(T, W00, c.P01), Y02 = fma(c2, W00, fma(T, c.P01, W02));
    const double Y11 = fma(T, c.P01, c.P11), Y12 = fma(c2, c.P01, fma(T, c.P11, c.P12));
    const double Y22 = fma(c2, W02, fma(T, c.P12, W22));
    const double X11 = fma(T, Y01, Y11), X12 = fma(T, Y02, Y12);
    const double X22 = fma(c2, Y02, fma(T, Y12, Y22));
    const double q0 = w0, q1 = fma(T, w0, w1), q2 = fma(c2, w0, fma(T, w1, w2));   // A'w
    // scale 2^k that brings pi' = sigma pi back to [1, 2): exponents only (exact)
    const double sc = pow2_from_biased(3069 - biased_exponent(sg) - biased_exponent(c.pi));
    const double ss = sg * sc, t0 = u0 * sc, t1 = u1 * sc, t2 = u2 * sc;
    // off the chain: true-scale gains  -K = -u / sigma,  -kc = -(B.w) / sigma
    const double is = -1.0 / sg;
    g01 = make_double2(u0 * is, u1 * is);
    g23 = make_double2(u2 * is, bw * is);
    c.ipi = -(c.ipi * is) * pow2_from_biased(2046 - biased_exponent(sc));    // 1 / (sigma sc pi)
    c.P00 = fma(ss, W00, -t0 * u0); c.P01 = fma(ss, Y01, -t0 * u1); c.P02 = fma(ss, Y02, -t0 * u2);
    c.P11 = fma(ss, X11, -t1 * u1); c.P12 = fma(ss, X12, -t1 * u2); c.P22 = fma(ss, X22, -t2 * u2);
    c.v0 = fma(ss, q0, -t0 * bw); c.v1 = fma(ss, q1, -t1 * bw); c.v2 = fma(ss, q2, -t2 * bw);
    c.pi = ss * c.pi;
}

// One PINNED stage (z_j = bnd fixes the jerk): returns the multiplier data P_{j+1} B/(CB), p_{j+1}.B/(CB).
__device__ __forceinline__ void backward_pinned(const RiccatiParams& P, CostToGo& c, double bnd, double2& g01, double2& g23) {
    const double hg = P.hg;
    const double m0 = fma(c.P02, P.bc[2], fma(c.P01, P.bc[1], c.P00 * P.bc[0]));
    const double m1 = fma(c.P12, P.bc[2], fma(c.P11, P.bc[1], c.P01 * P.bc[0]));
    const double m2 = fma(c.P22, P.bc[2], fma(c.P12, P.bc[1], c.P02 * P.bc[0]));
    const double mc = fma(c.v2, P.bc[2], fma(c.v1, P.bc[1], c.v0 * P.bc[0]));
    g01 = make_double2(m0 * c.ipi, m1 * c.ipi);
    g23 = make_double2(m2 * c.ipi, mc * c.ipi);
    const double u0 = fma(m0, bnd, c.v0), u1 = fma(m1, bnd, c.v1), u2 = fma(m2, bnd, c.v2);
    const double e2 = fma(hg, u0, u2);                 // Ac'u = a1 u1 + a2 (hg u0 + u2)
    const double prb = c.pi * (P.r_cb * bnd);
    c.v0 = fma(P.a2[0], e2, fma(P.a1[0], u1, prb * P.kv[0]));
    c.v1 = fma(P.a2[1], e2, fma(P.a1[1], u1, prb * P.kv[1]));
    c.v2 = fma(P.a2[2], e2, fma(P.a1[2], u1, prb * P.kv[2]));
    // M = L'P L (2x2),  P <- a~'M a~ + pi R kv kv'
    const double M11 = c.P11, M12 = fma(hg, c.P01, c.P12), M22 = fma(P.hgsq, c.P00, fma(P.hg2x, c.P02, c.P22));
    const double Y10 = fma(M12, P.a2[0], M11 * P.a1[0]), Y11 = fma(M12, P.a2[1], M11 * P.a1[1]), Y12 = fma(M12, P.a2[2], M11 * P.a1[2]);
    const double Y20 = fma(M22, P.a2[0], M12 * P.a1[0]), Y21 = fma(M22, P.a2[1], M12 * P.a1[1]), Y22 = fma(M22, P.a2[2], M12 * P.a1[2]);
    c.P00 = fma(P.a2[0], Y20, fma(P.a1[0], Y10, c.pi * P.rkk[0]));
    c.P01 = fma(P.a2[0], Y21, fma(P.a1[0], Y11, c.pi * P.rkk[1]));
    c.P02 = fma(P.a2[0], Y22, fma(P.a1[0], Y12, c.pi * P.rkk[2]));
    c.P11 = fma(P.a2[1], Y21, fma(P.a1[1], Y11, c.pi * P.rkk[3]));
    c.P12 = fma(P.a2[1], Y22, fma(P.a1[1], Y12, c.pi * P.rkk[4]));
    c.P22 = fma(P.a2[2], Y22, fma(P.a1[2], Y12, c.pi * P.rkk[5]));
}

// multiword bit helpers of the active-set masks (once per tick, not per stage)
__device__ __forceinline__ bool get_bit(const uint32_t* m, int j) { return (m[j >> 5] >> (j & 31)) & 1u; }

__device__ __forceinline__ int row_status(const uint32_t* lo, const uint32_t* hi, int j) {
    return get_bit(hi, j) ? 1 : (get_bit(lo, j) ? -1 : 0);
}
__device__ __forceinline__ void set_status(uint32_t* lo, uint32_t* hi, int j, int st) {
    const uint32_t bit = 1u << (j & 31);
    lo[j >> 5] = (lo[j >> 5] & ~bit) | (st < 0 ? bit : 0u);
    hi[j >> 5] = (hi[j >> 5] & ~bit) | (st > 0 ? bit : 0u);
}

// The guess of the next tick's active set, predicted from the last two (in: p = A_{k-1}, c = A_k, each in its own
// window coordinates; out: c = guess for tick k+1, p = A_k).  A boundary of the set (two neighbouring rows with
// different status) is either anchored in time -- it moves up by one row per tick with the window: the plain shift --
// or anchored to the window (terminal effects of a saturated plan: it stays where it is), or it moves with the state
// (a pinned run that keeps growing while the CoM falls).  Each boundary of A_k is matched with the same boundary of
// A_{k-1} (shift first, then stationary, then other small displacements) and moved on by the same amount; unmatched
// boundaries shift.  It is only a guess -- the active-set iteration corrects it -- but on the BASELINE push sweep it
// takes the heaviest column from 3.2 to 1.5 sweep pairs per tick.  Once per tick and column, not per stage.
__device__ __noinline__ void predict_active_set(uint32_t* p_lo, uint32_t* p_hi, uint32_t* c_lo, uint32_t* c_hi, int N, int NW) {
    uint32_t o_lo[kMaskWords], o_hi[kMaskWords];
    for (int w = 0; w < NW; ++w) {                 // default: plain shift, new last row free
        o_lo[w] = (c_lo[w] >> 1) | (((w + 1 < NW) ? c_lo[w + 1] : 0u) << 31);
        o_hi[w] = (c_hi[w] >> 1) | (((w + 1 < NW) ? c_hi[w + 1] : 0u) << 31);
    }
    for (int w = 0; w < NW; ++w) {
        // bit j set: rows j and j + 1 of A_k differ (the shifted copy holds row j + 1 at position j)
        uint32_t d = (c_lo[w] ^ o_lo[w]) | (c_hi[w] ^ o_hi[w]);
        if (w == NW - 1) d &= ~(1u << ((N - 1) & 31));        // the window end is handled below
        while (d) {
            const int q = w * 32 + __ffs((int)d) - 1;
            d &= d - 1u;
            const int a = row_status(c_lo, c_hi, q), b = row_status(c_lo, c_hi, q + 1);
            int v = -1;                                          // displacement per tick in window coordinates
            for (int k = 0; k < 6; ++k) {
                const int dq = k == 0 ? 1 : (k == 1 ? 0 : (k == 2 ? -1 : (k == 3 ? 2 : (k == 4 ? -2 : -3))));   // q' - q, by preference
                const int qp = q + dq;
                if (qp < 0 || qp + 1 >= N) continue;
                if (row_status(p_lo, p_hi, qp) == a && row_status(p_lo, p_hi, qp + 1) == b) { v = -dq; break; }
            }
            // the shift put this boundary at q - 1; move it to q + v
            for (int r = q; r <= q + v && r < N; ++r) if (r >= 0) set_status(o_lo, o_hi, r, a);
            for (int r = q + v + 1; r <= q - 1; ++r) if (r >= 0 && r < N) set_status(o_lo, o_hi, r, b);
        }
    }
    // the last row keeps its status when it had it on both ticks (anchored to the window end)
    {
        const int a = row_status(c_lo, c_hi, N - 1);
        if (a != 0 && row_status(p_lo, p_hi, N - 1) == a) set_status(o_lo, o_hi, N - 1, a);
    }
    for (int w = 0; w < NW; ++w) {
        p_lo[w] = c_lo[w]; p_hi[w] = c_hi[w];
        c_lo[w] = o_lo[w]; c_hi[w] = o_hi[w];
    }
}

// Bound series repacked per launch as (lo, hi) pairs: one 16-byte load per stage, 32-bit sample index.
// by_column = 0: one run of `ns` pairs per (series, axis); 1: one run per column of the launch (scenarios that bring
// their own series: the run of a column follows the processing order).
__global__ void k_pack_bounds(const RiccatiParams P, double2* __restrict__ out, int runs, int by_column) {
    const ColumnBatch& B = P.cb;
    const int ns = B.n_uniform;
    const long long total = (long long)runs * ns;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long run = e / ns;
        const int s = (int)(e - run * ns);
        long long bz;
        int axis;
        if (by_column) {
            const long long c = P.col_begin + run;
            if (c >= P.col_end) continue;
            axis = (int)(c / B.B);
            const int slot = (int)(c - (long long)axis * B.B);
            bz = B.order ? B.order[slot] : slot;
        } else {
            const int naxes = B.bound_axis_offset ? 2 : 1;
            bz = run / naxes;
            axis = (int)(run - bz * naxes);
        }
        const long long off = bz * B.bound_scn_stride + (long long)axis * B.bound_axis_offset + (long long)s * B.bound_stride;
        out[e] = make_double2(B.z_min[off], B.z_max[off]);
    }
}

// LANES = columns per warp (compile time: record strides become immediates).  SMEM: the per-plan stage tables and the
// feed-forward terms of the table stages live in shared memory (else: tables through L1, terms in global scratch).
// Registers: blocks of 64 threads.  Narrow warps (the resumed columns of a push sweep, small batches) are bound by the number
// of warps an SM holds -- a warp owns registers for 32 threads however few of them carry a column -- so they are capped at
// 200 registers (five blocks per SM, ~130 bytes of spills off the stage loops; push sweep of 1024 ... 8192 scenarios 3-5 %
// faster); full warps are bound by the FP64 pipe and keep all ~236 (four blocks per SM; the cap costs them 3-8 %).
#ifndef LIPM_RICCATI_NARROW_REGS
#define LIPM_RICCATI_NARROW_REGS 200
#endif
template <int LANES, bool SMEM>
__global__ void __maxnreg__(LANES <= 4 ? LIPM_RICCATI_NARROW_REGS : 255) k_riccati_rollout(const RiccatiParams P) {
    extern __shared__ double2 dyn_smem[];
    const ColumnBatch& B = P.cb;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int N = P.N;
    constexpr int KS = 2 * LANES;                 // doubles between consecutive stages in the feed-forward scratch (two warps per block)
    if (SMEM) {
        const double2* src = reinterpret_cast<const double2*>(P.tab);
        const int n2 = (N + 4) * 8;
        for (int k = threadIdx.x; k < n2; k += blockDim.x) dyn_smem[k] = __ldg(src + k);
        __syncthreads();
    }
    // per-plan stage tables, eight 16-byte words per stage
    const double2* const tab2 = SMEM ? dyn_smem : reinterpret_cast<const double2*>(P.tab);
    // feed-forward terms of the table stages: [stage][column of the block]
    double* const KC = (SMEM ? reinterpret_cast<double*>(dyn_smem + (N + 4) * 8) : P.nkc + (size_t)blockIdx.x * (N + 4) * KS) +
                       ((threadIdx.x >> 5) * LANES + lane);
    if (lane >= LANES) return;
    const long long cl = (long long)P.col_begin + warp * LANES + lane;
    if (cl >= P.col_end) return;
    const int c = (int)cl;
    if (P.park && P.park[c].stage != P.park_stage) return;      // resume mode: only the columns parked for this launch
    const int NW = (N + 31) >> 5;
    constexpr int GS = 2 * LANES;                 // double2 elements between consecutive stages of a column

    // ---- column set-up (same addressing as the ADMM kernels) ---------------------------------------------
    const int axis = c / B.B;
    const int slot = c - axis * B.B;
    const int b = B.order ? B.order[slot] : slot;
    int n = B.n_steps ? B.n_steps[b] : B.n_uniform;
    n = n > B.n_max ? B.n_max : n;
    int ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
    if (n < 1 || ticks < 0) ticks = 0;
    const long long bz = B.bounds_index ? B.bounds_index[b] : ((B.bounds_batch == 1) ? 0 : b);
    // (lo, hi) pairs of this column's series (k_pack_bounds)
    const double2* __restrict__ pb = P.pk + (P.pk_by_column ? (long long)(c - P.col_begin) : bz * (B.bound_axis_offset ? 2 : 1) + axis) * B.n_uniform;
    double* ph = axis ? B.hist[1] : B.hist[0];
    if (ph) ph += (long long)b * B.hist_scn_stride;
    uint8_t* pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
    double* pcom = B.com ? B.com + ((long long)b * B.n_max) * 2 + axis : nullptr;
    double* pzmp = B.zmp ? B.zmp + ((long long)b * B.n_max) * 2 + axis : nullptr;
    double* pplan = B.zmp_plan ? B.zmp_plan + (long long)c * N : nullptr;
    const double* s_init = (axis ? B.init[1] : B.init[0]) + (long long)b * 3;
    double x0 = s_init[0], x1 = s_init[1], x2 = s_init[2];
    const bool push = axis == 1 && B.impulse != nullptr;
    const double impulse = push ? B.impulse[b] : 0.0;
    const int ftick = push ? (B.force_tick ? B.force_tick[b] : -1) : -1;
    const int last = n - 1;
    double2* __restrict__ G = P.gains + ((size_t)warp * (N + 4) * LANES + lane) * 2;    // record j: G[j * GS], G[j * GS + 1]

    double viol = 0.0, lean = 0.0;
    int sweeps_total = 0, notconv = 0;
    long long general_stages = 0, table_stages = 0;     // statistics: stages that went through the general recursion / the tables
    if (B.write_initial && ticks > 0 && !P.park) {
        const double zmp = x0 - P.hg * x2;
        const double2 b0 = __ldg(pb);
        const double below = b0.x - zmp, above = zmp - b0.y;
        viol = fmax(fmax(below, above), 0.0);
        lean = fabs(x0 - zmp);
        if (ph) { ph[0] = x0; ph[1] = x1; ph[2] = x2; }
        if (pcom) pcom[0] = x0;
        if (pzmp) pzmp[0] = zmp;
        if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) | (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
    }

    // active set: bit j of m_lo / m_hi = row j pinned to its lower / upper bound; s_lo / s_hi = the final set of the
    // tick before (predict_active_set extrapolates from the two)
    uint32_t m_lo[kMaskWords], m_hi[kMaskWords], s_lo[kMaskWords], s_hi[kMaskWords];
#pragma unroll
    for (int w = 0; w < kMaskWords; ++w) m_lo[w] = m_hi[w] = s_lo[w] = s_hi[w] = 0u;

    const double T = P.T, c2 = P.c2, c3 = P.c3, hg = P.hg, Q = P.Q;
    const int N4 = N & ~3;                          // rows N4 .. N-1: a tail of up to three stages

    int i_begin = 0, calm = 0;
    bool handed_back = false;
    if (P.park) {
        // resume a column of the row-parallel kernel: its tick, state, running outcomes and current set
        const ParkRecord& rec = P.park[c];
        i_begin = rec.tick; sweeps_total = rec.sweeps; notconv = rec.notconv;
        x0 = rec.x[0]; x1 = rec.x[1]; x2 = rec.x[2]; viol = rec.viol; lean = rec.lean;
#pragma unroll
        for (int w = 0; w < kRowparWords; ++w) { m_lo[w] = s_lo[w] = rec.m_lo[w]; m_hi[w] = s_hi[w] = rec.m_hi[w]; }
    }
    for (int i = i_begin; i < ticks; ++i) {
        const int base = i + B.bound_index_bias;      // window row j reads sample min(base + j, last)
        // ---- guess: extrapolated from the last two ticks' sets (plain shift by default) --------------------
        if (i > i_begin) {
            predict_active_set(s_lo, s_hi, m_lo, m_hi, N, NW);
        }
        uint32_t zmag_hi = 0x3ff00000u;            // high word of max(1, |z|_inf) of the last forward sweep
        bool converged = false;
        int jg = 0;                                // first row of the table region in the last sweep
        bool schur = false;                        // the last sweep was a few-pin sweep (plan in KC, first sample in z_first)
        double z_first = 0.0;
        for (int sweep = 0; sweep < P.max_sweeps; ++sweep) {
            ++sweeps_total;
            // ---- leading pinned block: rows 0 .. m-1 are all pinned, so their jerks follow from the bounds alone and
            //      the cost-to-go is needed from stage m on only; their multipliers come from a costate pass -----
            int m = 0;
            for (int w = 0; w < NW; ++w) {
                const uint32_t freeb = ~(m_lo[w] | m_hi[w]);
                const int t = freeb ? (__ffs((int)freeb) - 1) : 32;
                m += t;
                if (t < 32) break;
            }
            m = m > N ? N : m;
            m &= ~3;                // the sweeps below work in aligned chunks of four stages (a shorter block is only less of a shortcut)
            // ---- table region: rows jg .. N-1 are free and so is every row after them ------------------------------
            int jt = -1;            // highest pinned row
            for (int w = NW - 1; w >= 0; --w) {
                const uint32_t pinned = m_lo[w] | m_hi[w];
                if (pinned) { jt = w * 32 + 31 - __clz((int)pinned); break; }
            }
            jg = (jt + 4) & ~3;
            if (jg > N4) jg = N;    // the general recursion takes the unaligned tail too
            // ---- few pinned rows (typically isolated rows where the plan cuts the corner of a support change): solve
            //      the unconstrained problem with the tables (rows 0 .. N-1), then correct it with the dense inverse
            //      Hessian:  z = z_unc - Hinv[:, A] lambda,  Hinv[A, A] lambda = z_unc[A] - b_A   (see few_pin_correction)
            int na = 0;
            for (int w = 0; w < NW; ++w) na += __popc(m_lo[w] | m_hi[w]);
            schur = na > 0 && na <= P.schur_max;
            if (schur) { m = 0; jg = 0; }
            const int jhi = jg == N ? N4 : jg;      // the general recursion's aligned chunks cover rows m .. jhi-1
            general_stages += (jg > m ? jg - m : 0); table_stages += N - jg;
            double s0 = x0, s1 = x1, s2 = x2;
            if (m > 0) {
                // chunks of four rows (m is a multiple of four); the bounds of a chunk are fetched while the chunk before it
                // is computed, the pin bits come one nibble per chunk
                double2 nb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) nb[k] = __ldg(pb + min(base + k, last));
                for (int j = 0; j < m; j += 4) {
                    double2 cb4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) cb4[k] = nb[k];
                    if (j + 4 < m) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nb[k] = __ldg(pb + min(base + j + 4 + k, last));
                    }
                    const uint32_t nib_hi = (m_hi[j >> 5] >> (j & 31)) & 15u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double lo = cb4[k].x, hi = cb4[k].y;
                        const double bnd = ((nib_hi >> k) & 1u) ? hi : lo;
                        const double Jv = fma(P.kv[2], s2, fma(P.kv[1], s1, fma(P.kv[0], s0, bnd * P.inv_cb)));
                        const double f0 = fma(c2, s2, fma(T, s1, s0)), f1 = fma(T, s2, s1);
                        s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                        G[(j + k) * GS] = make_double2(Jv, Q * (bnd - 0.5 * (lo + hi)));
                    }
                }
            }
            const double xm0 = s0, xm1 = s1, xm2 = s2;     // state at the start of stage m
            // ---- backward sweep, table region (rows N-1 .. jg): linear part of the cost-to-go only -----------------
            CostToGo ctg;
            ctg.P00 = ctg.P01 = ctg.P02 = ctg.P11 = ctg.P12 = ctg.P22 = ctg.v0 = ctg.v1 = ctg.v2 = 0.0;
            ctg.pi = ctg.ipi = 1.0;
            if (jg < N) {
                double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                auto table_backward = [&](int j, double lo, double hi) {
                    const double2* t = tab2 + j * 8;
                    const double2 a0 = t[0], a1 = t[1], b0 = t[2], b1 = t[3], e0 = t[4], e1 = t[5], d0 = t[6], d1 = t[7];
                    const double tq = P.q_half * (lo + hi);       // Q zref
                    const double n0 = fma(a0.x, v0, fma(a0.y, v1, fma(a1.x, v2, tq * a1.y)));
                    const double n1 = fma(b0.x, v0, fma(b0.y, v1, fma(b1.x, v2, tq * b1.y)));
                    const double n2 = fma(e0.x, v0, fma(e0.y, v1, fma(e1.x, v2, tq * e1.y)));
                    KC[j * KS] = fma(d0.x, v0, fma(d0.y, v1, fma(d1.x, v2, tq * d1.y)));
                    v0 = n0; v1 = n1; v2 = n2;
                };
                for (int j = N - 1; j >= N4; --j) {                      // tail
                    const double2 bd = __ldg(pb + min(base + j, last));
                    table_backward(j, bd.x, bd.y);
                }
                if (jg < N4) {
                    double blo[4], bhi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double2 bd = __ldg(pb + min(base + N4 - 4 + k, last));
                        blo[k] = bd.x; bhi[k] = bd.y;
                    }
                    for (int j = N4 - 4; j >= jg; j -= 4) {
                        double clo[4], chi[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) { clo[k] = blo[k]; chi[k] = bhi[k]; }
                        if (j > jg) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const double2 bd = __ldg(pb + min(base + j - 4 + k, last));
                                blo[k] = bd.x; bhi[k] = bd.y;
                            }
                        }
#pragma unroll
                        for (int k = 3; k >= 0; --k) table_backward(j + k, clo[k], chi[k]);
                    }
                }
                if (jg > 0) {       // hand the cost-to-go of stage jg to the general recursion / the costate pass
                    const double* a = P.aux + jg * 12;
                    ctg.P00 = __ldg(a + 4); ctg.P01 = __ldg(a + 5); ctg.P02 = __ldg(a + 6);
                    ctg.P11 = __ldg(a + 7); ctg.P12 = __ldg(a + 8); ctg.P22 = __ldg(a + 9);
                    ctg.v0 = v0; ctg.v1 = v1; ctg.v2 = v2;
                }
            }
            // ---- backward sweep, general recursion over rows jg-1 .. m ------------------------------------------------
            if (jg == N) {
                for (int j = N - 1; j >= N4 && j >= m; --j) {            // tail
                    const double2 bd = __ldg(pb + min(base + j, last));
                    const double lo = bd.x, hi = bd.y;
                    const bool pin_lo = get_bit(m_lo, j), pin_hi = get_bit(m_hi, j);
                    double2 g01, g23;
                    if (pin_lo | pin_hi) backward_pinned(P, ctg, pin_hi ? hi : lo, g01, g23);
                    else backward_free(P, ctg, lo, hi, g01, g23);
                    G[j * GS] = g01; G[j * GS + 1] = g23;
                }
            }
            if (m < jhi) {
                // chunks of four stages, highest first; the bounds of a chunk are fetched while the chunk before it is computed
                double blo[4], bhi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 bd = __ldg(pb + min(base + jhi - 4 + k, last));
                    blo[k] = bd.x; bhi[k] = bd.y;
                }
                double2* gj = G + (jhi - 4) * GS;
                for (int j = jhi - 4; j >= m; j -= 4) {
                    double clo[4], chi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { clo[k] = blo[k]; chi[k] = bhi[k]; }
                    if (j > m) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const double2 bd = __ldg(pb + min(base + j - 4 + k, last));
                            blo[k] = bd.x; bhi[k] = bd.y;
                        }
                    }
                    const uint32_t nib_lo = (m_lo[j >> 5] >> (j & 31)) & 15u, nib_hi = (m_hi[j >> 5] >> (j & 31)) & 15u;
#pragma unroll
                    for (int k = 3; k >= 0; --k) {
                        const bool pin_lo = (nib_lo >> k) & 1u, pin_hi = (nib_hi >> k) & 1u;
                        double2 g01, g23;
                        if (pin_lo | pin_hi) backward_pinned(P, ctg, pin_hi ? chi[k] : clo[k], g01, g23);
                        else backward_free(P, ctg, clo[k], chi[k], g01, g23);
                        gj[k * GS] = g01; gj[k * GS + 1] = g23;
                    }
                    gj -= 4 * GS;
                }
            }
            // ---- forward sweep over rows m .. N-1: plan, multipliers, corrected set ---------------------------
            bool changed = false;
            const double tol = P.mu_eps * Q * __hiloint2double((int)zmag_hi, 0);
            zmag_hi = 0x3ff00000u;
            // one stage of the forward sweep (returns the corrected pin bits of the row)
            auto forward_stage = [&](double lo, double hi, double2 ga, double2 gb, uint32_t pin_lo, uint32_t pin_hi,
                                     uint32_t& keep_lo, uint32_t& keep_hi) {
                const double f0 = fma(c2, s2, fma(T, s1, s0)), f1 = fma(T, s2, s1);     // A s, independent of the jerk
                double zi;
                keep_lo = pin_lo; keep_hi = pin_hi;
                if (pin_lo | pin_hi) {
                    const double bnd = pin_hi ? hi : lo;
                    const double Jv = fma(P.kv[2], s2, fma(P.kv[1], s1, fma(P.kv[0], s0, bnd * P.inv_cb)));
                    s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                    const double mu = fma(P.r_cb, Jv, fma(Q, bnd - 0.5 * (lo + hi), fma(ga.x, s0, fma(ga.y, s1, fma(gb.x, s2, gb.y)))));
                    if (pin_hi ? (mu > tol) : (mu < -tol)) { keep_lo = keep_hi = 0u; changed = true; }
                    zi = bnd;
                } else {
                    const double Jv = fma(ga.x, s0, fma(ga.y, s1, fma(gb.x, s2, gb.y)));
                    s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                    zi = fma(-hg, s2, s0);
                    if (zi > hi) { keep_hi = 1u; changed = true; }
                    else if (zi < lo) { keep_lo = 1u; changed = true; }
                }
                zmag_hi = max(zmag_hi, (uint32_t)__double2hiint(zi) & 0x7fffffffu);
            };
            // one stage of the table region: s <- Acl_j s - B kc_j, box test (returns 1: below, 2: above, 0: inside)
            auto table_forward = [&](int j, double lo, double hi, double nk) -> uint32_t {
                const double2* t = tab2 + j * 8;
                const double2 a0 = t[0], a1 = t[1], b0 = t[2], b1 = t[3], e0 = t[4], e1 = t[5];
                const double n0 = fma(a0.x, s0, fma(b0.x, s1, fma(e0.x, s2, c3 * nk)));
                const double n1 = fma(a0.y, s0, fma(b0.y, s1, fma(e0.y, s2, c2 * nk)));
                const double n2 = fma(a1.x, s0, fma(b1.x, s1, fma(e1.x, s2, T * nk)));
                s0 = n0; s1 = n1; s2 = n2;
                const double zi = fma(-hg, n2, n0);
                if (schur) { KC[j * KS] = zi; return 0u; }          // unconstrained plan; tested after the correction
                zmag_hi = max(zmag_hi, (uint32_t)__double2hiint(zi) & 0x7fffffffu);
                return zi > hi ? 2u : (zi < lo ? 1u : 0u);
            };
            if (m < jhi) {
                // chunks of four stages; gains and bounds of a chunk are fetched while the chunk before it is computed
                // (an L2 hit is ~250 cycles, a forward stage ~60), and prefetched into L1 kGainPrefetch stages ahead
                const double2* gj = G + m * GS;
                double2 na[4], nb[4];
                double blo[4], bhi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    na[k] = gj[k * GS]; nb[k] = gj[k * GS + 1];
                    const double2 bd = __ldg(pb + min(base + m + k, last));
                    blo[k] = bd.x; bhi[k] = bd.y;
                }
                for (int j = m; j < jhi; j += 4) {
                    double2 ga[4], gb[4];
                    double clo[4], chi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ga[k] = na[k]; gb[k] = nb[k]; clo[k] = blo[k]; chi[k] = bhi[k]; }
                    gj += 4 * GS;
                    if (j + 4 < jhi) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            na[k] = gj[k * GS]; nb[k] = gj[k * GS + 1];
                            const double2 bd = __ldg(pb + min(base + j + 4 + k, last));
                            blo[k] = bd.x; bhi[k] = bd.y;
                        }
                        prefetch_l1(gj + (kGainPrefetch - 4) * GS);
                        prefetch_l1(gj + (kGainPrefetch - 2) * GS);
                    }
                    const int w = j >> 5, sh = j & 31;
                    const uint32_t nib_lo = (m_lo[w] >> sh) & 15u, nib_hi = (m_hi[w] >> sh) & 15u;
                    uint32_t new_lo = 0u, new_hi = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t kl, kh;
                        forward_stage(clo[k], chi[k], ga[k], gb[k], (nib_lo >> k) & 1u, (nib_hi >> k) & 1u, kl, kh);
                        new_lo |= kl << k; new_hi |= kh << k;
                    }
                    if ((new_lo ^ nib_lo) | (new_hi ^ nib_hi)) {
                        m_lo[w] = (m_lo[w] & ~(15u << sh)) | (new_lo << sh);
                        m_hi[w] = (m_hi[w] & ~(15u << sh)) | (new_hi << sh);
                    }
                }
            }
            if (jg == N) {
                for (int j = (m > N4 ? m : N4); j < N; ++j) {              // tail
                    const double2 bd = __ldg(pb + min(base + j, last));
                    uint32_t kl, kh;
                    forward_stage(bd.x, bd.y, G[j * GS], G[j * GS + 1], get_bit(m_lo, j) ? 1u : 0u,
                                  get_bit(m_hi, j) ? 1u : 0u, kl, kh);
                    set_status(m_lo, m_hi, j, kh ? 1 : (kl ? -1 : 0));
                }
            } else {
                if (jg < N4) {
                    double nk[4], blo[4], bhi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        nk[k] = KC[(jg + k) * KS];
                        const double2 bd = __ldg(pb + min(base + jg + k, last));
                        blo[k] = bd.x; bhi[k] = bd.y;
                    }
                    for (int j = jg; j < N4; j += 4) {
                        double ck[4], clo[4], chi[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) { ck[k] = nk[k]; clo[k] = blo[k]; chi[k] = bhi[k]; }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {            // next chunk (the scratch has four stages of slack)
                            nk[k] = KC[(j + 4 + k) * KS];
                            const double2 bd = __ldg(pb + min(base + j + 4 + k, last));
                            blo[k] = bd.x; bhi[k] = bd.y;
                        }
                        if (!SMEM) prefetch_l1(KC + (j + kGainPrefetch) * KS);
                        uint32_t out = 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) out |= table_forward(j + k, clo[k], chi[k], ck[k]) << (2 * k);
                        if (out) {
                            changed = true;
                            const int w = j >> 5, sh = j & 31;
                            const uint32_t add_lo = (out & 1u) | ((out >> 1) & 2u) | ((out >> 2) & 4u) | ((out >> 3) & 8u);
                            const uint32_t add_hi = ((out >> 1) & 1u) | ((out >> 2) & 2u) | ((out >> 3) & 4u) | ((out >> 4) & 8u);
                            m_lo[w] |= add_lo << sh;
                            m_hi[w] |= add_hi << sh;
                        }
                    }
                }
                for (int j = N4; j < N; ++j) {                              // tail
                    const double2 bd = __ldg(pb + min(base + j, last));
                    const uint32_t r = table_forward(j, bd.x, bd.y, KC[j * KS]);
                    if (r) { changed = true; set_status(m_lo, m_hi, j, r == 2u ? 1 : -1); }
                }
            }
            // ---- few-pin sweep: dense correction of the unconstrained plan, box test, multipliers --------------------
            if (schur) {
                int idx[kSchurMax];
                double lam[kSchurMax], dinv[kSchurMax], S[kSchurMax * (kSchurMax + 1) / 2];
                int cnt = 0;
                for (int w = 0; w < NW; ++w) {
                    uint32_t bits = m_lo[w] | m_hi[w];
                    while (bits) {
                        const int j = w * 32 + __ffs((int)bits) - 1;
                        bits &= bits - 1u;
                        const double2 bd = __ldg(pb + min(base + j, last));
                        idx[cnt] = j;
                        lam[cnt] = KC[j * KS] - (get_bit(m_hi, j) ? bd.y : bd.x);       // z_unc - b
                        ++cnt;
                    }
                }
                // S = Hinv[A, A] (packed lower triangle), factorised as L D L' in place
                for (int k = 0; k < cnt; ++k) {
                    const double* row = P.hinv + (size_t)idx[k] * N;
                    for (int l = 0; l <= k; ++l) S[k * (k + 1) / 2 + l] = __ldg(row + idx[l]);
                }
                for (int k = 0; k < cnt; ++k) {
                    double* Lk = S + k * (k + 1) / 2;
                    for (int l = 0; l < k; ++l) {
                        const double* Ll = S + l * (l + 1) / 2;
                        double acc = Lk[l];
                        for (int q = 0; q < l; ++q) acc = fma(-Lk[q], Ll[q], acc);      // Lk[q] still holds L[k][q] d[q]
                        Lk[l] = acc;                                                     // = L[k][l] d[l]
                    }
                    double d = Lk[k];
                    for (int l = 0; l < k; ++l) {
                        const double ld = Lk[l];
                        const double lv = ld * dinv[l];
                        d = fma(-ld, lv, d);
                        Lk[l] = lv;
                    }
                    dinv[k] = 1.0 / d;
                }
                for (int k = 0; k < cnt; ++k) {                      // L y = r
                    const double* Lk = S + k * (k + 1) / 2;
                    double acc = lam[k];
                    for (int l = 0; l < k; ++l) acc = fma(-Lk[l], lam[l], acc);
                    lam[k] = acc;
                }
                for (int k = 0; k < cnt; ++k) lam[k] *= dinv[k];
                for (int k = cnt - 1; k >= 0; --k) {                 // L' lambda = y
                    double acc = lam[k];
                    for (int l = k + 1; l < cnt; ++l) acc = fma(-S[l * (l + 1) / 2 + k], lam[l], acc);
                    lam[k] = acc;
                }
                // corrected plan, four rows at a time; free rows are tested against the box
                zmag_hi = 0x3ff00000u;
                for (int j = 0; j < N; j += 4) {
                    double z[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) z[k] = KC[min(j + k, N - 1) * KS];
                    for (int q = 0; q < cnt; ++q) {
                        const double* row = P.hinv + (size_t)idx[q] * N + j;
                        const double lq = lam[q];
#pragma unroll
                        for (int k = 0; k < 4; ++k) z[k] = fma(-lq, __ldg(row + min(k, N - 1 - j)), z[k]);
                    }
                    const int w = j >> 5, sh = j & 31;
                    const uint32_t pinned = ((m_lo[w] | m_hi[w]) >> sh) & 15u;
                    uint32_t add_lo = 0u, add_hi = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (j + k < N) {
                            const double2 bd = __ldg(pb + min(base + j + k, last));
                            if ((pinned >> k) & 1u) {
                                z[k] = ((m_hi[w] >> (sh + k)) & 1u) ? bd.y : bd.x;
                            } else {
                                if (z[k] > bd.y) add_hi |= 1u << k;
                                else if (z[k] < bd.x) add_lo |= 1u << k;
                            }
                            zmag_hi = max(zmag_hi, (uint32_t)__double2hiint(z[k]) & 0x7fffffffu);
                            KC[(j + k) * KS] = z[k];
                        }
                    }
                    if (j == 0) z_first = z[0];
                    if (add_lo | add_hi) { changed = true; m_lo[w] |= add_lo << sh; m_hi[w] |= add_hi << sh; }
                }
                // multipliers: the gradient on a pinned row is -lambda; a row pinned on the wrong side is released
                for (int q = 0; q < cnt; ++q) {
                    const int j = idx[q];
                    const double mu = -lam[q];
                    if (get_bit(m_hi, j) ? (mu > tol) : (mu < -tol)) {
                        m_lo[j >> 5] &= ~(1u << (j & 31));
                        m_hi[j >> 5] &= ~(1u << (j & 31));
                        changed = true;
                    }
                }
            }
            // ---- costate pass over the leading block, rows m-1 .. 0: lambda_j = grad of the cost-to-go at x_j;
            //      mu_j = R J_j / (CB) + Q (b_j - zref_j) + B/(CB) . lambda_{j+1} ------------------------------------
            if (m > 0) {
                double l0 = fma(ctg.P02, xm2, fma(ctg.P01, xm1, fma(ctg.P00, xm0, ctg.v0))) * ctg.ipi;
                double l1 = fma(ctg.P12, xm2, fma(ctg.P11, xm1, fma(ctg.P01, xm0, ctg.v1))) * ctg.ipi;
                double l2 = fma(ctg.P22, xm2, fma(ctg.P12, xm1, fma(ctg.P02, xm0, ctg.v2))) * ctg.ipi;
                // chunks of four rows, highest first; the records of a chunk (written by the forward pass over the block: L2
                // hits) are fetched while the chunk before it is computed, the set is read and corrected one nibble per chunk
                double2 nr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) nr[k] = G[(m - 4 + k) * GS];
                for (int j = m - 4; j >= 0; j -= 4) {
                    double2 rec[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) rec[k] = nr[k];
                    if (j > 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nr[k] = G[(j - 4 + k) * GS];
                    }
                    const int w = j >> 5, sh = j & 31;
                    const uint32_t nib_hi = (m_hi[w] >> sh) & 15u;
                    uint32_t release = 0u;
#pragma unroll
                    for (int k = 3; k >= 0; --k) {
                        const double mu = fma(P.r_cb, rec[k].x, rec[k].y + fma(P.bc[2], l2, fma(P.bc[1], l1, P.bc[0] * l0)));
                        if (((nib_hi >> k) & 1u) ? (mu > tol) : (mu < -tol)) release |= 1u << k;
                        const double e2 = fma(hg, l0, l2), rj = P.R * rec[k].x;
                        const double t0 = fma(P.a2[0], e2, fma(P.a1[0], l1, rj * P.kv[0]));
                        const double t1 = fma(P.a2[1], e2, fma(P.a1[1], l1, rj * P.kv[1]));
                        const double t2 = fma(P.a2[2], e2, fma(P.a1[2], l1, rj * P.kv[2]));
                        l0 = t0; l1 = t1; l2 = t2;
                    }
                    if (release) {
                        m_lo[w] &= ~(release << sh);
                        m_hi[w] &= ~(release << sh);
                        changed = true;
                    }
                }
            }
            if (!changed) { converged = true; break; }
        }

        // ---- apply the first sample: state update, push, outputs ------------------------------------------
        {
            const double2 bd = __ldg(pb + min(base, last));
            const double l0 = bd.x, h0 = bd.y;
            const int first_pinned = (int)((m_lo[0] & 1u) | ((m_hi[0] & 1u) << 1));
            const double pred = fma(P.px02, x2, fma(T, x1, x0));      // Px[0] . x
            double z0;
            if (first_pinned) {
                z0 = (first_pinned & 2) ? h0 : l0;
            } else {
                double J0;
                if (schur) {                                     // few-pin sweep: the corrected plan is at hand
                    J0 = (z_first - pred) * P.inv_cb;
                } else if (jg > 0) {                             // row 0 went through the general recursion
                    const double2 ga = G[0], gb = G[1];
                    J0 = fma(ga.x, x0, fma(ga.y, x1, fma(gb.x, x2, gb.y)));
                } else {                                         // no pinned row at all: J = -K_0 . x - kc_0
                    J0 = fma(-__ldg(P.aux + 0), x0, fma(-__ldg(P.aux + 1), x1, fma(-__ldg(P.aux + 2), x2, KC[0])));
                }
                z0 = fma(P.cbv, J0, pred);
                z0 = z0 < l0 ? l0 : (z0 > h0 ? h0 : z0);             // a free row is inside the box; keep it there to the last bit
            }
            const double jerk = (z0 - pred) * P.inv_cb;
            double n0 = fma(c3, jerk, fma(c2, x2, fma(T, x1, x0)));
            double n1 = fma(c2, jerk, fma(T, x2, x1));
            double n2 = fma(T, jerk, x2);
            const double zmp = fma(-hg, n2, n0);
            const double below = l0 - zmp, above = zmp - h0;
            viol = fmax(viol, fmax(fmax(below, above), 0.0));
            lean = fmax(lean, fabs(n0 - zmp));
            unsigned f = (below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) | (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u);
            // on-the-bound bits by equality, like the ADMM kernels (a degenerate box sets both)
            f |= (z0 == l0 ? LIPM_FLAG_ACTIVE_LO : 0u) | (z0 == h0 ? LIPM_FLAG_ACTIVE_HI : 0u);
            if (!converged) {
                // like the ADMM kernels: the walk of this column ends at the first QP without a solution
                f |= LIPM_FLAG_NOT_CONVERGED;
                notconv += 1;
            }
            if (pplan && i == ticks - 1) {
                // the previewed ZMP of the last tick (single-tick API): one more forward pass over the final gains
                double s0 = x0, s1 = x1, s2 = x2;
                for (int j = 0; schur && j < N; ++j) pplan[j] = KC[j * KS];
                for (int j = 0; !schur && j < N; ++j) {
                    const double2 bd = __ldg(pb + min(base + j, last));
                    const bool plo_j = get_bit(m_lo, j), phi_j = get_bit(m_hi, j);
                    double Jv;
                    if (plo_j | phi_j) {
                        Jv = fma(P.kv[2], s2, fma(P.kv[1], s1, fma(P.kv[0], s0, (phi_j ? bd.y : bd.x) * P.inv_cb)));
                    } else if (j < jg) {
                        const double2 ga = G[j * GS], gb = G[j * GS + 1];
                        Jv = fma(ga.x, s0, fma(ga.y, s1, fma(gb.x, s2, gb.y)));
                    } else {
                        const double* a = P.aux + j * 12;
                        Jv = fma(-__ldg(a), s0, fma(-__ldg(a + 1), s1, fma(-__ldg(a + 2), s2, KC[j * KS])));
                    }
                    const double f0 = fma(c2, s2, fma(T, s1, s0)), f1 = fma(T, s2, s1);
                    s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                    pplan[j] = (plo_j | phi_j) ? (phi_j ? bd.y : bd.x) : fma(-hg, s2, s0);
                }
            }
            if (i == ftick) n1 -= impulse;
            const long long sample = i + 1;
            if (ph) {
                double* dst = ph + (sample - B.hist_index_bias) * 3;
                dst[0] = n0; dst[1] = n1; dst[2] = n2;
            }
            if (pf) pf[sample * 2] = (uint8_t)f;
            if (pcom) pcom[sample * 2] = n0;
            if (pzmp) pzmp[sample * 2] = zmp;
            x0 = n0; x1 = n1; x2 = n2;
            if (!converged) break;
        }
        // ---- resume mode: the push is absorbed (few pinned rows for a few ticks): back to the row-parallel kernel ---------
        if (P.back_list && i + 1 < ticks) {
            int pins = 0;
            for (int w = 0; w < NW; ++w) pins += __popc(m_lo[w] | m_hi[w]);
            calm = pins <= P.back_pins ? calm + 1 : 0;
            if (calm >= kBackStreak) {
                ParkRecord& rec = P.park[c];
                rec.tick = i + 1; rec.sweeps = sweeps_total; rec.notconv = notconv; rec.stage = P.back_stage;
                rec.x[0] = x0; rec.x[1] = x1; rec.x[2] = x2; rec.viol = viol; rec.lean = lean;
#pragma unroll
                for (int w = 0; w < kRowparWords; ++w) {        // the guess of the next tick: the set moved up by one row
                    rec.m_lo[w] = (m_lo[w] >> 1) | ((w + 1 < kMaskWords ? m_lo[w + 1] : 0u) << 31);
                    rec.m_hi[w] = (m_hi[w] >> 1) | ((w + 1 < kMaskWords ? m_hi[w + 1] : 0u) << 31);
                }
                P.back_list[atomicAdd(P.back_count, 1)] = c;
                if (P.stats) atomicAdd(P.stats + 3, 1ull);          // statistics: [3] columns handed back
                handed_back = true;
                break;
            }
        }
    }
    if (handed_back) return;

    if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
    if (B.lean_max) B.lean_max[b * 2 + axis] = lean;
    if (B.status) B.status[b * 2 + axis] = notconv;
    if (B.iters) B.iters[b * 2 + axis] = sweeps_total;
    if (P.stats) {
        atomicAdd(P.stats, (unsigned long long)sweeps_total);
        atomicAdd(P.stats + 1, 1ull);
        if (!P.park) {
            atomicAdd(P.stats + 2, (unsigned long long)general_stages);
            atomicAdd(P.stats + 3, (unsigned long long)table_stages);
        }
    }
}

template <int LANES, bool SMEM>
int launch_variant(const RiccatiParams& P, unsigned blocks, int threads, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        static bool granted[64] = {};   // per instantiation and device (the attribute belongs to the device's copy of the
        int dev = 0;                    // kernel); setting it twice from two threads is harmless
        LIPM_CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !granted[dev]) {
            LIPM_CUDA_TRY(cudaFuncSetAttribute(k_riccati_rollout<LANES, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynamicSmem));
            if (dev >= 0 && dev < 64) granted[dev] = true;
        }
    }
    k_riccati_rollout<LANES, SMEM><<<blocks, threads, smem, stream>>>(P);
    return LIPM_OK;
}

template <int LANES>
int launch_lanes(const RiccatiParams& P, unsigned blocks, int threads, size_t smem, cudaStream_t stream) {
    return smem ? launch_variant<LANES, true>(P, blocks, threads, smem, stream) : launch_variant<LANES, false>(P, blocks, threads, 0, stream);
}

int launch_kernel(lipm_plan* plan, const RiccatiParams& P0, cudaStream_t stream) {
    RiccatiParams P = P0;
    const int ncols = P.cb.ncols;
    if (ncols <= 0) return LIPM_OK;
    // columns per warp: give every scheduler (4 per SM) a warp before widening the warps
    int lanes = plan->riccati_lanes;
    if (lanes <= 0) {
        const long long want_warps = (long long)plan->sm_count * 8;       // ~2 warps per scheduler before the warps widen (measured: 4096 scenarios run best on 8 columns per warp)
        lanes = 1;
        while (lanes < 32 && (long long)ncols > want_warps * lanes) lanes <<= 1;
    }
    // resume: a third of the columns is left, in runs of the processing order, and the columns that share a warp no longer
    // follow the same path through the stage types (measured best at 1024 ... 8192 scenarios: a quarter of the width)
    if (P.park && plan->riccati_lanes <= 0) lanes = lanes >= 4 ? lanes >> 2 : 1;
    P.lanes = lanes;
    LaunchState state;
    if (!P.stats) {                 // (resume: the statistics words of the row-parallel stage are continued)
        LIPM_CUDA_TRY(state.begin(stream));
        P.stats = state.stats();
    }
    const int warps_per_block = 2;
    // shared memory: the per-plan stage tables (128 B per stage) and the feed-forward terms of the block's table stages
    // (8 B per stage and column).  When that is too much for three blocks per SM (long horizons, wide warps) the tables
    // are read through L1 and the terms go to global scratch.
    const size_t smem_bytes = (size_t)(P.N + 4) * (128 + 8 * warps_per_block * lanes);
    const bool in_smem = plan->riccati_scratch == 1 || (plan->riccati_scratch == 0 && smem_bytes <= kMaxAutoSmem);
    if (in_smem && smem_bytes > kMaxDynamicSmem) {
        set_error("active-set kernel: %zu bytes of shared memory requested (horizon %d, %d columns per warp)", smem_bytes, P.N, lanes);
        return LIPM_E_UNSUPPORTED;
    }
    const size_t smem = in_smem ? smem_bytes : 0;
    // bound series as (lo, hi) pairs: per (series, axis) when the series are shared or indexed, per column otherwise
    const ColumnBatch& cb = P.cb;
    const int ns = cb.n_uniform;
    P.pk_by_column = (!cb.bounds_index && cb.bounds_batch != 1) ? 1 : 0;
    // scratch per column: feedback data of the general recursion (32 B per stage), the table stages' feed-forward when it
    // is not in shared memory (8 B per stage), the column's own bound pairs; batches beyond 4 GiB run in several launches
    const size_t per_col = (size_t)(P.N + 4) * (in_smem ? 32 : 40) + (P.pk_by_column ? (size_t)ns * 16 : 0);
    long long cols_per_launch = (long long)((size_t)4 << 30) / (long long)per_col;
    const int cols_per_block = warps_per_block * lanes;
    cols_per_launch -= cols_per_launch % cols_per_block;
    if (cols_per_launch < cols_per_block) cols_per_launch = cols_per_block;
    if (cols_per_launch > ncols) cols_per_launch = ((long long)ncols + cols_per_block - 1) / cols_per_block * cols_per_block;
    const size_t gain_bytes = (size_t)cols_per_launch * (P.N + 4) * 32 + (size_t)(kGainPrefetch + 8) * 32 * 32;
    const size_t kc_bytes = in_smem ? 0 : (size_t)cols_per_launch * (P.N + 4) * 8 + (size_t)(kGainPrefetch + 8) * 64 * 8;
    const long long pk_runs = P.pk_by_column ? cols_per_launch : (long long)cb.bounds_batch * (cb.bound_axis_offset ? 2 : 1);
    const size_t pk_bytes = (size_t)pk_runs * ns * 16;
    StreamBuffer scratch;
    LIPM_CUDA_TRY(scratch.alloc(gain_bytes + kc_bytes + pk_bytes, stream));
    P.gains = scratch.as<double2>();
    P.nkc = in_smem ? nullptr : reinterpret_cast<double*>(scratch.as<unsigned char>() + gain_bytes);
    double2* pk = reinterpret_cast<double2*>(scratch.as<unsigned char>() + gain_bytes + kc_bytes);
    P.pk = pk;
    for (long long c0 = 0; c0 < ncols; c0 += cols_per_launch) {
        P.col_begin = (int)c0;
        P.col_end = (int)(c0 + cols_per_launch < ncols ? c0 + cols_per_launch : ncols);
        if (P.pk_by_column || c0 == 0) {
            const long long runs = P.pk_by_column ? (long long)(P.col_end - P.col_begin) : pk_runs;
            const long long total = runs * ns;
            const unsigned pblocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)plan->sm_count * 16);
            k_pack_bounds<<<pblocks ? pblocks : 1, 256, 0, stream>>>(P, pk, (int)runs, P.pk_by_column);
            LIPM_CUDA_TRY(cudaGetLastError());
        }
        const long long warps = ((long long)(P.col_end - P.col_begin) + lanes - 1) / lanes;
        const long long blocks = (warps + warps_per_block - 1) / warps_per_block;
        int rc;
        switch (lanes) {
            case 1: rc = launch_lanes<1>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
            case 2: rc = launch_lanes<2>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
            case 4: rc = launch_lanes<4>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
            case 8: rc = launch_lanes<8>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
            case 16: rc = launch_lanes<16>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
            default: rc = launch_lanes<32>(P, (unsigned)blocks, 32 * warps_per_block, smem, stream); break;
        }
        if (rc != LIPM_OK) return rc;
        LIPM_CUDA_TRY(cudaGetLastError());
    }
    if (!P.park) LIPM_CUDA_TRY(cudaMemcpyAsync(plan->d_stats, P.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    return LIPM_OK;        // (resume mode: the caller files the statistics of the whole sequence)
}

}  // namespace

// Per-plan tables of the stages whose later rows are all free (host, once per plan).  Stage j, j = N-1 .. 0, with
// W = P_{j+1} + Q C'C:  sigma = R + B'W B,  K = B'W A / sigma,  Acl = A - B K,  P_j = Acl'W Acl + R K'K  (Joseph form).
// tab[j] = four groups of four: group g < 3 = (Acl[0][g], Acl[1][g], Acl[2][g], -(Acl'C')_g), group 3 = (-B / sigma, C B / sigma),
// so that  v_j = sum_r group_g[r] v_{j+1,r} + Q zref group_g[3]  and  -kc_j = group_3 . (v_{j+1}, Q zref).
int build_riccati_tables(lipm_plan* plan) {
    const lipm_plan_desc_t& d = plan->desc;
    const int N = plan->N;
    const double T = d.dt, c2 = d.dt * d.dt / 2.0, c3 = d.dt * d.dt * d.dt / 6.0, hg = d.h / d.g;
    const double A[3][3] = {{1.0, T, c2}, {0.0, 1.0, T}, {0.0, 0.0, 1.0}};
    const double Bv[3] = {c3, c2, T};
    const double Cv[3] = {1.0, 0.0, -hg};
    std::vector<double> tab((size_t)(N + 4) * 16, 0.0), aux((size_t)(N + 1) * 12, 0.0);
    double Pm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = N - 1; j >= 0; --j) {
        double W[3][3], WB[3], K[3], Acl[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) W[r][c] = Pm[r][c] + d.Q * Cv[r] * Cv[c];
        double sigma = d.R;
        for (int r = 0; r < 3; ++r) {
            WB[r] = W[r][0] * Bv[0] + W[r][1] * Bv[1] + W[r][2] * Bv[2];
        }
        for (int r = 0; r < 3; ++r) sigma += Bv[r] * WB[r];
        for (int c = 0; c < 3; ++c) K[c] = (A[0][c] * WB[0] + A[1][c] * WB[1] + A[2][c] * WB[2]) / sigma;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Acl[r][c] = A[r][c] - Bv[r] * K[c];
        double WA[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) WA[r][c] = W[r][0] * Acl[0][c] + W[r][1] * Acl[1][c] + W[r][2] * Acl[2][c];
        double Pn[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                Pn[r][c] = Acl[0][r] * WA[0][c] + Acl[1][r] * WA[1][c] + Acl[2][r] * WA[2][c] + d.R * K[r] * K[c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Pm[r][c] = 0.5 * (Pn[r][c] + Pn[c][r]);
        double* t = &tab[(size_t)j * 16];
        for (int g = 0; g < 3; ++g) {
            t[4 * g + 0] = Acl[0][g]; t[4 * g + 1] = Acl[1][g]; t[4 * g + 2] = Acl[2][g];
            t[4 * g + 3] = -(Acl[0][g] * Cv[0] + Acl[1][g] * Cv[1] + Acl[2][g] * Cv[2]);
        }
        t[12] = -Bv[0] / sigma; t[13] = -Bv[1] / sigma; t[14] = -Bv[2] / sigma;
        t[15] = (Bv[0] * Cv[0] + Bv[1] * Cv[1] + Bv[2] * Cv[2]) / sigma;
        double* a = &aux[(size_t)j * 12];
        a[0] = K[0]; a[1] = K[1]; a[2] = K[2];
        a[4] = Pm[0][0]; a[5] = Pm[0][1]; a[6] = Pm[0][2]; a[7] = Pm[1][1]; a[8] = Pm[1][2]; a[9] = Pm[2][2];
    }
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_ric_tab, tab.size() * sizeof(double)));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_ric_aux, aux.size() * sizeof(double)));
    LIPM_CUDA_TRY(cudaMemcpy(plan->d_ric_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    LIPM_CUDA_TRY(cudaMemcpy(plan->d_ric_aux, aux.data(), aux.size() * sizeof(double), cudaMemcpyHostToDevice));
    return LIPM_OK;
}

static int launch_riccati_stage(lipm_plan* plan, const ColumnBatch& cb, const RowparScratch* rs, int chunk, int in_list, int back_list,
                                cudaStream_t stream);

namespace {
// side streams / events of one chunked rollout
struct SideStreams {
    cudaStream_t stream[kRowparMaxChunks] = {};
    cudaEvent_t ready[kRowparMaxChunks] = {}, done[kRowparMaxChunks] = {};
    int count = 0;
    cudaError_t create(int n) {
        for (int k = 0; k < n; ++k) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
            count = k + 1;
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~SideStreams() {
        for (int k = 0; k < count; ++k) {
            if (ready[k]) cudaEventDestroy(ready[k]);
            if (done[k]) cudaEventDestroy(done[k]);
            if (stream[k]) cudaStreamDestroy(stream[k]);
        }
    }
};

// Side-stream sets are kept per device and handed out for the duration of one call's ENQUEUE (creating and destroying four
// streams and eight events per rollout is a dozen driver calls on the host path of a 38 ms step, and it showed with eight
// ranks on one host).  A set may be reused as soon as the call that held it has returned: the waits of that call were
// enqueued with the events' state at that time, and work on a reused stream simply queues behind the earlier call's.
// The sets live until the process ends (a host application that resets the device must not call the library afterwards
// with plans created before the reset -- the same rule as for the plans themselves).
class SidePool {
public:
    SideStreams* acquire(int device, cudaError_t* err) {
        *err = cudaSuccess;
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto& idle = idle_[device & 63];
            if (!idle.empty()) {
                SideStreams* s = idle.back().release();
                idle.pop_back();
                return s;
            }
        }
        std::unique_ptr<SideStreams> fresh(new SideStreams());
        *err = fresh->create(kRowparMaxChunks);
        return *err == cudaSuccess ? fresh.release() : nullptr;
    }
    void release(int device, SideStreams* s) {
        std::lock_guard<std::mutex> lock(mu_);
        idle_[device & 63].emplace_back(s);
    }
private:
    std::mutex mu_;
    std::vector<std::unique_ptr<SideStreams>> idle_[64];
};
SidePool& side_pool() {
    static SidePool* pool = new SidePool();       // never destroyed: no CUDA calls from static destructors at exit
    return *pool;
}
struct SideLease {
    int device = 0;
    SideStreams* set = nullptr;
    ~SideLease() { if (set) side_pool().release(device, set); }
};
}  // namespace

// Default FP64 strict solver.  Walks on shared bound series start on the row-parallel kernel (lipm_rollout_rowpar.cu); a
// column whose active set outgrows it continues here, and goes back once the set is small again:
//   row-parallel (all columns) -> Riccati (parked, hands back) -> row-parallel (handed back) -> Riccati (the rest).
// The Riccati stage is bound by the latency of its longest walks and leaves most of the machine idle, the row-parallel
// stage by throughput.  Batches of 4096 scenarios or more are therefore cut into chunks of the processing order (early
// pushes first: the walks that spend longest on the Riccati kernel): the row-parallel passes of the chunks run one
// after the other on the caller's stream, and each chunk's remaining three launches follow on a side stream as soon as
// its first pass is done, beside the first passes of the later chunks.
// Every other batch (own series per scenario, single ticks, horizons beyond 160) runs here from the first tick.
int launch_strict_rollout_riccati(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    if (plan->N > kMaxStrictHorizon) {
        set_error("strict mode: horizon %d exceeds this build's limit %d", plan->N, kMaxStrictHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    if (!rowpar_eligible(plan, cb)) return launch_riccati_stage(plan, cb, nullptr, 0, -1, -1, stream);
    RowparScratch rs;
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    rs.stats = state.stats();
    int rc = rowpar_prepare(plan, cb, rs, stream);
    if (rc != LIPM_OK) return rc;
    // (walks without a push never leave the row-parallel kernel for long: one pass over the whole batch is fastest, 8.0 against
    // 12.7 ms for 4096 scenarios, since every chunk's pass is as long as one walk)
    const bool pushed = cb.impulse != nullptr && cb.force_tick != nullptr;
    const int chunks = plan->rowpar_chunks > 0 ? (plan->rowpar_chunks < cb.B ? plan->rowpar_chunks : 1)
                                               : (cb.B >= 4096 && pushed ? kRowparAutoChunks : 1);
    if (chunks == 1) {
        if ((rc = launch_rowpar(plan, cb, rs, 0, 1, -1, 0, stream)) != LIPM_OK) return rc;
        if ((rc = launch_riccati_stage(plan, cb, &rs, 0, 0, 1, stream)) != LIPM_OK) return rc;
        if ((rc = launch_rowpar(plan, cb, rs, 0, 1, 1, 2, stream)) != LIPM_OK) return rc;
        rc = launch_riccati_stage(plan, cb, &rs, 0, 2, -1, stream);
    } else {
        SideLease lease;
        LIPM_CUDA_TRY(cudaGetDevice(&lease.device));
        {
            cudaError_t err = cudaSuccess;
            lease.set = side_pool().acquire(lease.device, &err);
            LIPM_CUDA_TRY(err);
        }
        SideStreams& side = *lease.set;
        // (no early return between here and the join: a failed call sets rc and leaves the loop)
        auto cuda_ok = [&rc](cudaError_t e, const char* what) {
            if (e == cudaSuccess) return true;
            set_error("%s failed: %s (chunked active-set rollout)", what, cudaGetErrorString(e));
            rc = LIPM_E_CUDA;
            return false;
        };
        for (int k = 0; k < chunks && rc == LIPM_OK; ++k) {
            if ((rc = launch_rowpar(plan, cb, rs, k, chunks, -1, 0, stream)) != LIPM_OK) break;
            if (!cuda_ok(cudaEventRecord(side.ready[k], stream), "cudaEventRecord")) break;
            if (!cuda_ok(cudaStreamWaitEvent(side.stream[k], side.ready[k], 0), "cudaStreamWaitEvent")) break;
            if ((rc = launch_riccati_stage(plan, cb, &rs, k, 0, 1, side.stream[k])) != LIPM_OK) break;
            if ((rc = launch_rowpar(plan, cb, rs, k, chunks, 1, 2, side.stream[k])) != LIPM_OK) break;
            if ((rc = launch_riccati_stage(plan, cb, &rs, k, 2, -1, side.stream[k])) != LIPM_OK) break;
            if (!cuda_ok(cudaEventRecord(side.done[k], side.stream[k]), "cudaEventRecord")) break;
        }
        // join: whatever was enqueued on the side streams precedes everything the caller enqueues next (and the release of
        // the scratch buffers below)
        // (after a failed launch the events of the later chunks carry no record of this call: waiting on a pooled event's
        // earlier, completed record -- or on none -- returns at once)
        for (int k = 0; k < chunks; ++k) cudaStreamWaitEvent(stream, side.done[k], 0);
        if (rc != LIPM_OK) {
            for (int k = 0; k < chunks; ++k) cudaStreamSynchronize(side.stream[k]);     // error path: nothing may outlive the scratch
        }
    }
    if (rc != LIPM_OK) return rc;
    LIPM_CUDA_TRY(cudaMemcpyAsync(plan->d_stats, rs.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    return LIPM_OK;
}

static int launch_riccati_stage(lipm_plan* plan, const ColumnBatch& cb, const RowparScratch* rs, int chunk, int in_list, int back_list,
                                cudaStream_t stream) {
    const lipm_plan_desc_t& d = plan->desc;
    if (plan->N > kMaxStrictHorizon) {
        set_error("strict mode: horizon %d exceeds this build's limit %d", plan->N, kMaxStrictHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    RiccatiParams P;
    P.cb = cb;
    P.N = plan->N;
    P.max_sweeps = plan->riccati_max_sweeps > 0 ? plan->riccati_max_sweeps : 64;
    P.lanes = 32;
    P.Q = d.Q; P.R = d.R;
    P.T = plan->T; P.c2 = std::pow(d.dt, 2) / 2.0; P.c3 = std::pow(d.dt, 3) / 6.0; P.hg = plan->hg;
    P.cbv = plan->pu00; P.inv_cb = 1.0 / plan->pu00;
    P.px02 = P.c2 - P.hg;
    P.r_cb = d.R / plan->pu00;
    P.q_half = 0.5 * d.Q;
    P.q_hg = d.Q * P.hg; P.q_hg2 = d.Q * P.hg * P.hg;
    P.s_const = d.R + d.Q * plan->pu00 * plan->pu00;
    const double Bv[3] = {P.c3, P.c2, P.T};
    P.bb[0] = Bv[0] * Bv[0]; P.bb[1] = 2.0 * Bv[0] * Bv[1]; P.bb[2] = 2.0 * Bv[0] * Bv[2];
    P.bb[3] = Bv[1] * Bv[1]; P.bb[4] = 2.0 * Bv[1] * Bv[2]; P.bb[5] = Bv[2] * Bv[2];
    const double A[9] = {1.0, P.T, P.c2, 0.0, 1.0, P.T, 0.0, 0.0, 1.0};
    const double ca[3] = {1.0, P.T, P.px02};
    for (int r = 0; r < 3; ++r) {
        P.kv[r] = -ca[r] / plan->pu00;
        P.bc[r] = Bv[r] / plan->pu00;
        P.a1[r] = A[3 + r] - Bv[1] * ca[r] / plan->pu00;
        P.a2[r] = A[6 + r] - Bv[2] * ca[r] / plan->pu00;
    }
    P.hg2x = 2.0 * P.hg; P.hgsq = P.hg * P.hg;
    P.col_begin = 0; P.col_end = cb.ncols; P.gains = nullptr; P.nkc = nullptr;
    P.tab = plan->d_ric_tab; P.aux = plan->d_ric_aux; P.pk = nullptr; P.pk_by_column = 0;
    P.park = nullptr; P.park_stage = 0; P.back_list = nullptr; P.back_count = nullptr; P.back_pins = 0; P.back_stage = 0;
    P.hinv = plan->d_hinv; P.schur_max = plan->riccati_few_pins <= 0 ? 0 : (plan->riccati_few_pins > kSchurMax ? kSchurMax : plan->riccati_few_pins);
    P.rkk[0] = d.R * P.kv[0] * P.kv[0]; P.rkk[1] = d.R * P.kv[0] * P.kv[1]; P.rkk[2] = d.R * P.kv[0] * P.kv[2];
    P.rkk[3] = d.R * P.kv[1] * P.kv[1]; P.rkk[4] = d.R * P.kv[1] * P.kv[2]; P.rkk[5] = d.R * P.kv[2] * P.kv[2];
    P.viol_tol = d.viol_tol;
    P.mu_eps = 1e-12;
    P.stats = nullptr;
    if (rs) {
        P.park = rs->park; P.park_stage = RowparScratch::stage(chunk, in_list); P.stats = rs->stats;
        if (back_list >= 0 && LIPM_BACK_MARGIN < kRowparMaxPins) {
            P.back_list = rs->list(chunk, back_list); P.back_count = rs->count(chunk, back_list);
            P.back_pins = kRowparMaxPins - LIPM_BACK_MARGIN;      // some room for the set to grow again before the row-parallel kernel would park it
            P.back_stage = RowparScratch::stage(chunk, back_list);
        }
    }
    return launch_kernel(plan, P, stream);
}

}  // namespace lipm
