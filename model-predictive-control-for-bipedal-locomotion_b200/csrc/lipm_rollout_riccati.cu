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
// Determinism: a column's arithmetic does not depend on its neighbours, so results are bit-identical whatever
// batch, warp shape or GPU a scenario lands in.
#include "lipm_internal.cuh"

#include <cmath>

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
    double2* gains;         // [warp][stage][lane] x 2: per-stage feedback data of the backward sweep
    unsigned long long* stats;   // [0] sweep pairs, [1] columns
};

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 2^k as a double from a biased exponent field (no range checks: callers keep it inside [1, 2046])
__device__ __forceinline__ double pow2_from_biased(int e) { return __hiloint2double(e << 20, 0); }
__device__ __forceinline__ int biased_exponent(double x) { return (__double2hiint(x) >> 20) & 0x7ff; }

constexpr int kGainPrefetch = 16;  // stages of look-ahead of the forward sweep's L1 prefetch (a forward stage is ~20 instructions, an L2 hit ~250 cycles)
constexpr int kMaskWords = kMaxStrictHorizon / 32;

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
__device__ __forceinline__ void set_bit(uint32_t* m, int j) { m[j >> 5] |= 1u << (j & 31); }

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

// LANES = columns per warp (compile time: record strides become immediates).
template <int LANES>
__global__ void __launch_bounds__(64, 4) k_riccati_rollout(const RiccatiParams P) {
    const ColumnBatch& B = P.cb;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lane >= LANES) return;
    const long long cl = (long long)P.col_begin + warp * LANES + lane;
    if (cl >= P.col_end) return;
    const int c = (int)cl;
    const int N = P.N;
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
    const double* __restrict__ plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
    const double* __restrict__ phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
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
    const int bstride = B.bound_stride;
    double2* __restrict__ G = P.gains + ((size_t)warp * N * LANES + lane) * 2;    // record j: G[j * GS], G[j * GS + 1]

    double viol = 0.0, lean = 0.0;
    int sweeps_total = 0, notconv = 0;
    if (B.write_initial && ticks > 0) {
        const double zmp = x0 - P.hg * x2;
        const double below = plo[0] - zmp, above = zmp - phi[0];
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

    for (int i = 0; i < ticks; ++i) {
        const int base = i + B.bound_index_bias;      // window row j reads sample min(base + j, last)
        // ---- guess: extrapolated from the last two ticks' sets (plain shift by default) --------------------
        if (i > 0) {
            predict_active_set(s_lo, s_hi, m_lo, m_hi, N, NW);
        }
        uint32_t zmag_hi = 0x3ff00000u;            // high word of max(1, |z|_inf) of the last forward sweep
        bool converged = false;
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
            double s0 = x0, s1 = x1, s2 = x2;
            for (int j = 0; j < m; ++j) {
                const int off = min(base + j, last) * bstride;
                const double lo = __ldg(plo + off), hi = __ldg(phi + off);
                const double bnd = get_bit(m_hi, j) ? hi : lo;
                const double Jv = fma(P.kv[2], s2, fma(P.kv[1], s1, fma(P.kv[0], s0, bnd * P.inv_cb)));
                const double f0 = fma(c2, s2, fma(T, s1, s0)), f1 = fma(T, s2, s1);
                s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                G[j * GS] = make_double2(Jv, Q * (bnd - 0.5 * (lo + hi)));
            }
            const double xm0 = s0, xm1 = s1, xm2 = s2;     // state at the start of stage m
            const int N4 = N & ~3;                          // rows N4 .. N-1: a tail of up to three stages
            // ---- backward sweep over rows N-1 .. m --------------------------------------------------------------
            CostToGo ctg;
            ctg.P00 = ctg.P01 = ctg.P02 = ctg.P11 = ctg.P12 = ctg.P22 = ctg.v0 = ctg.v1 = ctg.v2 = 0.0;
            ctg.pi = ctg.ipi = 1.0;
            for (int j = N - 1; j >= N4 && j >= m; --j) {            // tail
                const int off = min(base + j, last) * bstride;
                const double lo = __ldg(plo + off), hi = __ldg(phi + off);
                const bool pin_lo = get_bit(m_lo, j), pin_hi = get_bit(m_hi, j);
                double2 g01, g23;
                if (pin_lo | pin_hi) backward_pinned(P, ctg, pin_hi ? hi : lo, g01, g23);
                else backward_free(P, ctg, lo, hi, g01, g23);
                G[j * GS] = g01; G[j * GS + 1] = g23;
            }
            if (m < N4) {
                // chunks of four stages, highest first; the bounds of a chunk are fetched while the chunk before it is computed
                double blo[4], bhi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int off = min(base + N4 - 4 + k, last) * bstride;
                    blo[k] = __ldg(plo + off); bhi[k] = __ldg(phi + off);
                }
                double2* gj = G + (N4 - 4) * GS;
                for (int j = N4 - 4; j >= m; j -= 4) {
                    double clo[4], chi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { clo[k] = blo[k]; chi[k] = bhi[k]; }
                    if (j > m) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int off = min(base + j - 4 + k, last) * bstride;
                            blo[k] = __ldg(plo + off); bhi[k] = __ldg(phi + off);
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
            if (m < N4) {
                // chunks of four stages; gains and bounds of a chunk are fetched while the chunk before it is computed
                // (an L2 hit is ~250 cycles, a forward stage ~60), and prefetched into L1 kGainPrefetch stages ahead
                const double2* gj = G + m * GS;
                double2 na[4], nb[4];
                double blo[4], bhi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    na[k] = gj[k * GS]; nb[k] = gj[k * GS + 1];
                    const int off = min(base + m + k, last) * bstride;
                    blo[k] = __ldg(plo + off); bhi[k] = __ldg(phi + off);
                }
                for (int j = m; j < N4; j += 4) {
                    double2 ga[4], gb[4];
                    double clo[4], chi[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ga[k] = na[k]; gb[k] = nb[k]; clo[k] = blo[k]; chi[k] = bhi[k]; }
                    gj += 4 * GS;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {            // next chunk (the scratch has slack past the last stage)
                        na[k] = gj[k * GS]; nb[k] = gj[k * GS + 1];
                        const int off = min(base + j + 4 + k, last) * bstride;
                        blo[k] = __ldg(plo + off); bhi[k] = __ldg(phi + off);
                    }
                    prefetch_l1(gj + (kGainPrefetch - 4) * GS);
                    prefetch_l1(gj + (kGainPrefetch - 2) * GS);
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
            for (int j = (m > N4 ? m : N4); j < N; ++j) {              // tail
                const int off = min(base + j, last) * bstride;
                uint32_t kl, kh;
                forward_stage(__ldg(plo + off), __ldg(phi + off), G[j * GS], G[j * GS + 1], get_bit(m_lo, j) ? 1u : 0u,
                              get_bit(m_hi, j) ? 1u : 0u, kl, kh);
                set_status(m_lo, m_hi, j, kh ? 1 : (kl ? -1 : 0));
            }
            // ---- costate pass over the leading block, rows m-1 .. 0: lambda_j = grad of the cost-to-go at x_j;
            //      mu_j = R J_j / (CB) + Q (b_j - zref_j) + B/(CB) . lambda_{j+1} ------------------------------------
            if (m > 0) {
                double l0 = fma(ctg.P02, xm2, fma(ctg.P01, xm1, fma(ctg.P00, xm0, ctg.v0))) * ctg.ipi;
                double l1 = fma(ctg.P12, xm2, fma(ctg.P11, xm1, fma(ctg.P01, xm0, ctg.v1))) * ctg.ipi;
                double l2 = fma(ctg.P22, xm2, fma(ctg.P12, xm1, fma(ctg.P02, xm0, ctg.v2))) * ctg.ipi;
                for (int j = m - 1; j >= 0; --j) {
                    const double2 rec = G[j * GS];
                    const double mu = fma(P.r_cb, rec.x, rec.y + fma(P.bc[2], l2, fma(P.bc[1], l1, P.bc[0] * l0)));
                    const bool pin_hi = get_bit(m_hi, j);
                    if (pin_hi ? (mu > tol) : (mu < -tol)) {
                        m_lo[j >> 5] &= ~(1u << (j & 31));
                        m_hi[j >> 5] &= ~(1u << (j & 31));
                        changed = true;
                    }
                    const double e2 = fma(hg, l0, l2), rj = P.R * rec.x;
                    const double t0 = fma(P.a2[0], e2, fma(P.a1[0], l1, rj * P.kv[0]));
                    const double t1 = fma(P.a2[1], e2, fma(P.a1[1], l1, rj * P.kv[1]));
                    const double t2 = fma(P.a2[2], e2, fma(P.a1[2], l1, rj * P.kv[2]));
                    l0 = t0; l1 = t1; l2 = t2;
                }
            }
            if (!changed) { converged = true; break; }
        }

        // ---- apply the first sample: state update, push, outputs ------------------------------------------
        {
            const int off = min(base, last) * bstride;
            const double l0 = __ldg(plo + off), h0 = __ldg(phi + off);
            const int first_pinned = (int)((m_lo[0] & 1u) | ((m_hi[0] & 1u) << 1));
            const double pred = fma(P.px02, x2, fma(T, x1, x0));      // Px[0] . x
            double z0;
            if (first_pinned) {
                z0 = (first_pinned & 2) ? h0 : l0;
            } else {
                const double2 ga = G[0], gb = G[1];
                const double J0 = fma(ga.x, x0, fma(ga.y, x1, fma(gb.x, x2, gb.y)));
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
                for (int j = 0; j < N; ++j) {
                    const int o = min(base + j, last) * bstride;
                    const double2 ga = G[j * GS], gb = G[j * GS + 1];
                    const bool plo_j = get_bit(m_lo, j), phi_j = get_bit(m_hi, j);
                    double Jv;
                    if (plo_j | phi_j) Jv = fma(P.kv[2], s2, fma(P.kv[1], s1, fma(P.kv[0], s0, (phi_j ? __ldg(phi + o) : __ldg(plo + o)) * P.inv_cb)));
                    else Jv = fma(ga.x, s0, fma(ga.y, s1, fma(gb.x, s2, gb.y)));
                    const double f0 = fma(c2, s2, fma(T, s1, s0)), f1 = fma(T, s2, s1);
                    s0 = fma(c3, Jv, f0); s1 = fma(c2, Jv, f1); s2 = fma(T, Jv, s2);
                    pplan[j] = (plo_j | phi_j) ? (phi_j ? __ldg(phi + o) : __ldg(plo + o)) : fma(-hg, s2, s0);
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
    }

    if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
    if (B.lean_max) B.lean_max[b * 2 + axis] = lean;
    if (B.status) B.status[b * 2 + axis] = notconv;
    if (B.iters) B.iters[b * 2 + axis] = sweeps_total;
    if (P.stats) {
        atomicAdd(P.stats, (unsigned long long)sweeps_total);
        atomicAdd(P.stats + 1, 1ull);
    }
}

template <int LANES>
void launch_lanes(const RiccatiParams& P, unsigned blocks, int threads, cudaStream_t stream) {
    k_riccati_rollout<LANES><<<blocks, threads, 0, stream>>>(P);
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
    P.lanes = lanes;
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    P.stats = state.stats();
    // per-stage feedback data: 32 bytes per stage and column; batches beyond 4 GiB of it run in several launches
    const size_t per_col = (size_t)P.N * 32;
    long long cols_per_launch = (long long)((size_t)4 << 30) / (long long)per_col;
    cols_per_launch -= cols_per_launch % lanes;
    if (cols_per_launch < lanes) cols_per_launch = lanes;
    if (cols_per_launch > ncols) cols_per_launch = ((long long)ncols + lanes - 1) / lanes * lanes;
    StreamBuffer scratch;
    // (+ kGainPrefetch stages of slack: the forward sweep looks ahead without range checks)
    LIPM_CUDA_TRY(scratch.alloc((size_t)cols_per_launch * per_col + (size_t)(kGainPrefetch + 8) * 32 * 32, stream));
    P.gains = scratch.as<double2>();
    const int warps_per_block = 2;
    for (long long c0 = 0; c0 < ncols; c0 += cols_per_launch) {
        P.col_begin = (int)c0;
        P.col_end = (int)(c0 + cols_per_launch < ncols ? c0 + cols_per_launch : ncols);
        const long long warps = ((long long)(P.col_end - P.col_begin) + lanes - 1) / lanes;
        const long long blocks = (warps + warps_per_block - 1) / warps_per_block;
        switch (lanes) {
            case 1: launch_lanes<1>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
            case 2: launch_lanes<2>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
            case 4: launch_lanes<4>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
            case 8: launch_lanes<8>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
            case 16: launch_lanes<16>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
            default: launch_lanes<32>(P, (unsigned)blocks, 32 * warps_per_block, stream); break;
        }
        LIPM_CUDA_TRY(cudaGetLastError());
    }
    LIPM_CUDA_TRY(state.publish(plan->d_stats));
    return LIPM_OK;
}

}  // namespace

int launch_strict_rollout_riccati(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
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
    P.col_begin = 0; P.col_end = cb.ncols; P.gains = nullptr;
    P.rkk[0] = d.R * P.kv[0] * P.kv[0]; P.rkk[1] = d.R * P.kv[0] * P.kv[1]; P.rkk[2] = d.R * P.kv[0] * P.kv[2];
    P.rkk[3] = d.R * P.kv[1] * P.kv[1]; P.rkk[4] = d.R * P.kv[1] * P.kv[2]; P.rkk[5] = d.R * P.kv[2] * P.kv[2];
    P.viol_tol = d.viol_tol;
    P.mu_eps = 1e-12;
    P.stats = nullptr;
    return launch_kernel(plan, P, stream);
}

}  // namespace lipm
