/*
 * lipm_oracle.c -- plain-C CPU restatement of the reference's Wieber path (TEST INFRASTRUCTURE).
 *
 * Checker and CPU baseline only: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product never does.
 *
 * Follows /root/reference/src/mpc_bipedal/controllers/zmp_controller.py:
 *   model A, B, C ........................ :15-21
 *   Px, Pu ................................ :162-171
 *   non-strict closed form ................ :196-199  (first jerk only, as the rollout uses it)
 *   strict QP ............................. :173-195  (unique minimiser; see below)
 *   rollout, bound padding, push .......... :59-108
 *
 * Strict path: **parity unpinned** by the reference (arithmetic inside cvxpy/OSQP, un-pinned, not
 * installed, no golden vectors).  The QP is strictly convex and always feasible, so its minimiser
 * is unique; this file computes it in ZMP coordinates z = Px x + Pu J,
 *     min Q/2 |z - zref|^2 + R/2 |Pu^-1 (z - Px x)|^2   s.t.  zmin <= z <= zmax,
 * by fixed-rho ADMM with over-relaxation run to `eps`.  oracle/lipm_oracle.py holds an
 * independent active-set solver (SciPy BVLS) that pins this one (tests/test_oracle.py).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared).  Scenarios are independent; the
 * rollout entry point parallelises over them with OpenMP.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int N;
    double dt, h, g, Q, R, rho, alpha, eps, viol_tol;
    int max_iter;
    double A[9], B[3], C[3];
    double *Px;   /* [N,3] */
    double *Pu;   /* [N,N] */
    double *M;    /* [N,N] (Q I + G + rho I)^-1 */
    double *Hinv; /* [N,N] (Q I + G)^-1 */
    double *GPx;  /* [N,3] */
    double *gain; /* [N+3] */
} oracle_plan;

static int cholesky_inverse(int n, double *W /* in: SPD, out: destroyed */, double *out) {
    for (int k = 0; k < n; ++k) {
        double d = W[k * n + k];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        W[k * n + k] = d;
        for (int i = k + 1; i < n; ++i) W[i * n + k] /= d;
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j <= i; ++j) W[i * n + j] -= W[i * n + k] * W[j * n + k];
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < j; ++i) out[i * n + j] = 0.0;
        for (int i = j; i < n; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) s -= W[i * n + k] * out[k * n + j];
            out[i * n + j] = s / W[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = out[i * n + j];
            for (int k = i + 1; k < n; ++k) s -= W[k * n + i] * out[k * n + j];
            out[i * n + j] = s / W[i * n + i];
        }
    }
    return 0;
}

void lipm_oracle_plan_free(oracle_plan *p) {
    if (!p) return;
    free(p->Px); free(p->Pu); free(p->M); free(p->Hinv); free(p->GPx); free(p->gain);
    free(p);
}

/* fp-contract off: Px and Pu must round exactly like the reference's float64 expressions */
__attribute__((optimize("fp-contract=off")))
oracle_plan *lipm_oracle_plan_create(int N, double h, double g, double dt, double Q, double R, double rho,
                                     double alpha, double eps, int max_iter, double viol_tol) {
    oracle_plan *p = (oracle_plan *)calloc(1, sizeof(oracle_plan));
    const double T = dt;
    p->N = N; p->dt = dt; p->h = h; p->g = g; p->Q = Q; p->R = R;
    p->rho = rho > 0 ? rho : 2.0;
    p->alpha = alpha > 0 ? alpha : 1.75;
    p->eps = eps > 0 ? eps : 1e-12;
    p->max_iter = max_iter > 0 ? max_iter : 2000;
    p->viol_tol = viol_tol >= 0 ? viol_tol : 1e-9;
    const double c3 = pow(T, 3) / 6.0, c2 = pow(T, 2) / 2.0, thg = T * h / g, hg = h / g;
    const double A[9] = {1.0, T, c2, 0.0, 1.0, T, 0.0, 0.0, 1.0};
    memcpy(p->A, A, sizeof(A));
    p->B[0] = c3; p->B[1] = c2; p->B[2] = T;
    p->C[0] = 1.0; p->C[1] = 0.0; p->C[2] = -hg;
    const size_t nn = (size_t)N * N;
    p->Px = (double *)malloc(sizeof(double) * N * 3);
    p->Pu = (double *)calloc(nn, sizeof(double));
    p->M = (double *)malloc(sizeof(double) * nn);
    p->Hinv = (double *)malloc(sizeof(double) * nn);
    p->GPx = (double *)malloc(sizeof(double) * N * 3);
    p->gain = (double *)malloc(sizeof(double) * (N + 3));
    for (int i = 0; i < N; ++i) {                       /* zmp_controller.py:162-171 */
        p->Px[i * 3 + 0] = 1.0;
        p->Px[i * 3 + 1] = T * (i + 1);
        p->Px[i * 3 + 2] = c2 * ((double)(i + 1) * (i + 1)) - hg;
        for (int j = 0; j <= i; ++j) {
            const double d = (double)(i - j);
            p->Pu[(size_t)i * N + j] = c3 * (1 + 3 * d + 3 * d * d) - thg;
        }
    }
    /* first column of Pu^-1 (Pu is lower-triangular Toeplitz) */
    double *t = (double *)malloc(sizeof(double) * N);
    t[0] = 1.0 / p->Pu[0];
    for (int k = 1; k < N; ++k) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += p->Pu[(size_t)(k - j) * N] * t[j];
        t[k] = -s / p->Pu[0];
    }
    double *G = (double *)malloc(sizeof(double) * nn);
    double *W = (double *)malloc(sizeof(double) * nn);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int k = (i > j ? i : j); k < N; ++k) s += t[k - i] * t[k - j];
            G[(size_t)i * N + j] = R * s;
        }
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += G[(size_t)i * N + j] * p->Px[j * 3 + c];
            p->GPx[i * 3 + c] = s;
        }
    int bad = 0;
    memcpy(W, G, sizeof(double) * nn);
    for (int i = 0; i < N; ++i) W[(size_t)i * N + i] += Q;
    bad |= cholesky_inverse(N, W, p->Hinv);
    memcpy(W, G, sizeof(double) * nn);
    for (int i = 0; i < N; ++i) W[(size_t)i * N + i] += Q + p->rho;
    bad |= cholesky_inverse(N, W, p->M);
    const double pu00 = p->Pu[0];
    for (int j = 0; j < N; ++j) p->gain[j] = Q * p->Hinv[j] / pu00;
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += p->Hinv[j] * p->GPx[j * 3 + c];
        p->gain[N + c] = (s - p->Px[c]) / pu00;
    }
    free(t); free(G); free(W);
    if (bad) { lipm_oracle_plan_free(p); return NULL; }
    return p;
}

void lipm_oracle_plan_export(const oracle_plan *p, double *Px, double *Pu, double *M, double *Hinv, double *gain) {
    const size_t nn = (size_t)p->N * p->N;
    if (Px) memcpy(Px, p->Px, sizeof(double) * p->N * 3);
    if (Pu) memcpy(Pu, p->Pu, sizeof(double) * nn);
    if (M) memcpy(M, p->M, sizeof(double) * nn);
    if (Hinv) memcpy(Hinv, p->Hinv, sizeof(double) * nn);
    if (gain) memcpy(gain, p->gain, sizeof(double) * (p->N + 3));
}

static inline double clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

/* One strict solve.  tv: in = warm start (relaxed iterate + scaled dual), out = final.
 * Returns iterations executed; *converged says whether the residual test passed. */
static int admm_solve(const oracle_plan *p, const double *x, const double *lo, const double *hi, double *tv,
                      double *rhs, double *q, int *converged) {
    const int N = p->N;
    const double rho = p->rho, alpha = p->alpha;
    for (int j = 0; j < N; ++j) {
        const double gx = p->GPx[j * 3] * x[0] + p->GPx[j * 3 + 1] * x[1] + p->GPx[j * 3 + 2] * x[2];
        q[j] = -(0.5 * p->Q * (lo[j] + hi[j]) + gx);
        const double z = clip(tv[j], lo[j], hi[j]);
        rhs[j] = rho * (2.0 * z - tv[j]) - q[j];          /* rho (z - w) - q,  w = t - z */
    }
    /* stop test on d = |v - z_old|: primal residual |v - z_new| <= (1+alpha) d and dual residual
     * rho |z_new - z_old| <= rho alpha d, so d <= thr bounds both by eps * max(1, |q|_inf)
     * (relative to the linear term: a diverging "fall" has a huge q and an absolute test would
     * sit below FP64 rounding) */
    double qmag = 0.0;
    for (int j = 0; j < N; ++j) qmag = fmax(qmag, fabs(q[j]));
    const double thr = p->eps * fmax(1.0, qmag) / fmax(1.0 + alpha, rho * alpha);
    *converged = 0;
    int it = 0;
    while (it < p->max_iter) {
        ++it;
        double dmax = 0.0;
        for (int i = 0; i < N; ++i) {
            const double *Mi = p->M + (size_t)i * N;
            double v = 0.0;
            for (int j = 0; j < N; ++j) v += Mi[j] * rhs[j];
            const double zo = clip(tv[i], lo[i], hi[i]);
            const double d = v - zo;
            dmax = fmax(dmax, fabs(d));
            tv[i] = tv[i] + alpha * d;                    /* alpha v + (1-alpha) zo + (t - zo) */
        }
        for (int i = 0; i < N; ++i) rhs[i] = rho * (2.0 * clip(tv[i], lo[i], hi[i]) - tv[i]) - q[i];
        if (dmax <= thr) { *converged = 1; break; }
    }
    return it;
}

/* ------------------------------------------------------------------------------------------------
 * Third strict solver (cross-check of the ADMM above and blueprint of csrc/lipm_rollout_riccati.cu):
 * primal-dual active-set iteration whose equality-constrained sub-problems are solved EXACTLY in O(N)
 * by a Riccati sweep.  The strict QP (zmp_controller.py:173-191) is an LQ tracking problem of the
 * 3-state model (:15-21): stage i applies jerk J_i, x_{i+1} = A x_i + B J_i, previews z_i = C x_{i+1},
 * costs Q/2 (z_i - zref_i)^2 + R/2 J_i^2 and must keep zmin_i <= z_i <= zmax_i.  For a guessed active set
 * (rows pinned to a bound) the backward sweep gives the cost-to-go 1/2 x'P_i x + p_i'x; a pinned row fixes its own
 * jerk (J_i = (b_i - C A x_i) / (C B)).  The forward sweep returns the plan z and, on pinned rows, the
 * multiplier mu_i = d cost / d b_i (= the gradient (H z + q)_i of the condensed QP).  Rows that leave the
 * box are pinned, pinned rows whose multiplier has the wrong sign are released, until nothing changes:
 * that fixed point satisfies the KKT conditions of the reference's QP, whose minimiser is unique.
 * act: in = guess (previous tick's set moved up by one sample), out = final set (-1 lower, +1 upper).
 * Returns Riccati sweeps executed; *converged = 0 if the set still changed after max_sweeps. */
/* Prediction of the next tick's active set from the last two.  A boundary of the set (two neighbouring rows with
 * different status) is either anchored in time -- it moves up by one row per tick with the window: the plain shift --
 * or anchored to the window (terminal effects of a saturated plan: it stays), or it moves with the state (a pinned run
 * that keeps growing).  Each boundary of `cur` is matched with the same boundary of `prev` (shift first, then
 * stationary, then +-1, +-2 rows) and moved on by the same amount; unmatched boundaries shift.  Only a guess: the
 * active-set iteration corrects it. */
static void predict_active_set(int N, const int8_t *prev, const int8_t *cur, int8_t *out) {
    for (int j = 0; j + 1 < N; ++j) out[j] = cur[j + 1];          /* default: plain shift, new last row free */
    out[N - 1] = 0;
    static const int cand[6] = {1, 0, -1, 2, -2, -3};             /* q' - q, in order of preference */
    for (int q = 0; q < N; ++q) {
        const int a = cur[q], b = (q + 1 < N) ? cur[q + 1] : 2;    /* 2 = beyond the window */
        if (a == b) continue;
        int v = -1;
        for (int k = 0; k < 6; ++k) {
            const int qp = q + cand[k];
            if (qp < 0 || qp >= N) continue;
            const int pa = prev[qp], pb = (qp + 1 < N) ? prev[qp + 1] : 2;
            if (pa == a && pb == b) { v = -cand[k]; break; }
        }
        if (b == 2 && v != 0) continue;                            /* the window end does not move */
        /* the shift put this boundary at q - 1; move it to q + v: rows q .. q + v take the lower side's status */
        for (int r = q; r <= q + v && r < N; ++r) if (r >= 0) out[r] = (int8_t)a;
        /* (v < -1: rows q + v + 1 .. q - 1 take the upper side's status) */
        if (b != 2) for (int r = q + v + 1; r <= q - 1; ++r) if (r >= 0 && r < N) out[r] = (int8_t)b;
    }
}

static uint8_t *g_sweep_log = NULL;   /* debug: [2B, n_max] sweeps per tick (tools only) */
void lipm_oracle_set_sweep_log(uint8_t *p) { g_sweep_log = p; }

static int riccati_pdas_solve(const oracle_plan *p, const double *x, const double *lo, const double *hi, int8_t *act,
                              double *gain /* [N,4] */, double *z, double *jerk0, int max_sweeps, int *converged) {
    const int N = p->N;
    const double T = p->dt, c2 = p->B[1], c3 = p->B[0], hg = -p->C[2], Q = p->Q, R = p->R;
    const double cb = c3 - hg * T;                      /* C B = Pu[0][0] */
    const double ca[3] = {1.0, T, c2 - hg};             /* C A = Px[0]    */
    const double kv[3] = {-ca[0] / cb, -ca[1] / cb, -ca[2] / cb};
    const double bc[3] = {c3 / cb, c2 / cb, T / cb};
    double Ac[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ac[r * 3 + c] = p->A[r * 3 + c] - p->B[r] * ca[c] / cb;
    *converged = 0;
    int sweeps = 0;
    double mu_scale = 1.0;
    while (sweeps < max_sweeps) {
        ++sweeps;
        /* backward: P symmetric (P00 P01 P02 P11 P12 P22), p */
        double P00 = 0, P01 = 0, P02 = 0, P11 = 0, P12 = 0, P22 = 0, p0 = 0, p1 = 0, p2 = 0;
        for (int i = N - 1; i >= 0; --i) {
            double *gi = gain + (size_t)i * 4;
            if (act[i]) {
                const double b = act[i] > 0 ? hi[i] : lo[i];
                const double m0 = P00 * bc[0] + P01 * bc[1] + P02 * bc[2];
                const double m1 = P01 * bc[0] + P11 * bc[1] + P12 * bc[2];
                const double m2 = P02 * bc[0] + P12 * bc[1] + P22 * bc[2];
                gi[0] = m0; gi[1] = m1; gi[2] = m2; gi[3] = bc[0] * p0 + bc[1] * p1 + bc[2] * p2;
                const double v0 = m0 * b + p0, v1 = m1 * b + p1, v2 = m2 * b + p2;
                const double rb = R * b / cb;
                const double n0 = Ac[0] * v0 + Ac[3] * v1 + Ac[6] * v2 + rb * kv[0];
                const double n1 = Ac[1] * v0 + Ac[4] * v1 + Ac[7] * v2 + rb * kv[1];
                const double n2 = Ac[2] * v0 + Ac[5] * v1 + Ac[8] * v2 + rb * kv[2];
                /* Y = P Ac, Pn = Ac' Y + R kv kv' */
                double Y[9];
                const double Pm[9] = {P00, P01, P02, P01, P11, P12, P02, P12, P22};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) Y[r * 3 + c] = Pm[r * 3] * Ac[c] + Pm[r * 3 + 1] * Ac[3 + c] + Pm[r * 3 + 2] * Ac[6 + c];
#define ACTY(r, c) (Ac[r] * Y[c] + Ac[3 + r] * Y[3 + c] + Ac[6 + r] * Y[6 + c] + R * kv[r] * kv[c])
                P00 = ACTY(0, 0); P01 = ACTY(0, 1); P02 = ACTY(0, 2); P11 = ACTY(1, 1); P12 = ACTY(1, 2); P22 = ACTY(2, 2);
#undef ACTY
                p0 = n0; p1 = n1; p2 = n2;
            } else {
                const double r = 0.5 * (lo[i] + hi[i]);
                const double W00 = P00 + Q, W01 = P01, W02 = P02 - Q * hg, W11 = P11, W12 = P12, W22 = P22 + Q * hg * hg;
                const double w0 = p0 - Q * r, w1 = p1, w2 = p2 + Q * r * hg;
                const double b0 = W00 * c3 + W01 * c2 + W02 * T;
                const double b1 = W01 * c3 + W11 * c2 + W12 * T;
                const double b2 = W02 * c3 + W12 * c2 + W22 * T;
                const double s = R + c3 * b0 + c2 * b1 + T * b2;
                const double bw = c3 * w0 + c2 * w1 + T * w2;
                const double u0 = b0, u1 = T * b0 + b1, u2 = c2 * b0 + T * b1 + b2;
                const double is = 1.0 / s;
                const double k0 = u0 * is, k1 = u1 * is, k2 = u2 * is, kc = bw * is;
                gi[0] = k0; gi[1] = k1; gi[2] = k2; gi[3] = kc;
                /* X = A' W A */
                const double Y00 = W00, Y01 = T * W00 + W01, Y02 = c2 * W00 + T * W01 + W02;
                const double Y11 = T * W01 + W11, Y12 = c2 * W01 + T * W11 + W12;
                const double Y22 = c2 * W02 + T * W12 + W22;
                const double X00 = Y00, X01 = Y01, X02 = Y02;
                const double X11 = T * Y01 + Y11, X12 = T * Y02 + Y12;
                const double X22 = c2 * Y02 + T * Y12 + Y22;
                P00 = X00 - u0 * k0; P01 = X01 - u0 * k1; P02 = X02 - u0 * k2;
                P11 = X11 - u1 * k1; P12 = X12 - u1 * k2; P22 = X22 - u2 * k2;
                p0 = w0 - u0 * kc; p1 = T * w0 + w1 - u1 * kc; p2 = c2 * w0 + T * w1 + w2 - u2 * kc;
            }
        }
        /* forward: plan, multipliers, set update */
        double s0 = x[0], s1 = x[1], s2 = x[2];
        int changed = 0;
        double zmag = 1.0;
        const double tol = 1e-12 * mu_scale;
        for (int i = 0; i < N; ++i) {
            const double *gi = gain + (size_t)i * 4;
            double J, zi;
            int a = act[i];
            if (a) {
                const double b = a > 0 ? hi[i] : lo[i];
                J = kv[0] * s0 + kv[1] * s1 + kv[2] * s2 + b / cb;
                const double n0 = s0 + T * s1 + c2 * s2 + c3 * J, n1 = s1 + T * s2 + c2 * J, n2 = s2 + T * J;
                const double mu = R * J / cb + Q * (b - 0.5 * (lo[i] + hi[i])) + gi[0] * n0 + gi[1] * n1 + gi[2] * n2 + gi[3];
                if ((a > 0 && mu > tol) || (a < 0 && mu < -tol)) { a = 0; changed = 1; }
                zi = b;
                s0 = n0; s1 = n1; s2 = n2;
            } else {
                J = -(gi[0] * s0 + gi[1] * s1 + gi[2] * s2 + gi[3]);
                const double n0 = s0 + T * s1 + c2 * s2 + c3 * J, n1 = s1 + T * s2 + c2 * J, n2 = s2 + T * J;
                zi = n0 - hg * n2;
                if (zi > hi[i]) { a = 1; changed = 1; }
                else if (zi < lo[i]) { a = -1; changed = 1; }
                s0 = n0; s1 = n1; s2 = n2;
            }
            if (i == 0) *jerk0 = J;
            z[i] = zi;
            zmag = fmax(zmag, fabs(zi));
            act[i] = (int8_t)a;
        }
        mu_scale = Q * zmag;
        if (!changed) { *converged = 1; break; }
    }
    return sweeps;
}

/*
 * Walks of B scenarios.  Same argument meaning as lipm_rollout in include/lipm_mpc.h, with host
 * pointers.  mode 0 = non-strict (zmp_controller.py:196-199), 1 = strict (ADMM), 2 = strict (active set + Riccati;
 * iters then counts Riccati sweeps).  Returns 0.
 */
int lipm_oracle_rollout(const oracle_plan *p, int mode, int B, int n_max, const int32_t *n_steps, const double *x0,
                        const double *y0, const double *z_min, const double *z_max, int bounds_batch,
                        const double *impulse, const int32_t *force_tick, double *x_hist, double *y_hist,
                        uint8_t *flags, double *max_violation, int32_t *status, int32_t *iters, int threads) {
    const int N = p->N;
    const double T = p->dt, c2 = p->B[1], c3 = p->B[0], hg = -p->C[2];
    const double px02 = p->Px[2], pu00 = p->Pu[0];
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
    {
        double *tv = (double *)malloc(sizeof(double) * N * 10);
        double *lo = tv + N, *hi = tv + 2 * N, *rhs = tv + 3 * N, *q = tv + 4 * N, *gain = tv + 5 * N, *zplan = tv + 9 * N;
        int8_t *act = (int8_t *)malloc(3 * (size_t)N);
        int8_t *older = act + N, *guess = act + 2 * N;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int col = 0; col < 2 * B; ++col) {
            const int axis = col / B, b = col % B;
            const int n = n_steps ? (n_steps[b] < n_max ? n_steps[b] : n_max) : n_max;
            const size_t bz = bounds_batch == 1 ? 0 : (size_t)b;
            const double *plo = z_min + bz * n_max * 2 + axis;
            const double *phi = z_max + bz * n_max * 2 + axis;
            double *hist = (axis ? y_hist : x_hist);
            if (hist) hist += (size_t)b * n_max * 3;
            uint8_t *pf = flags ? flags + (size_t)b * n_max * 2 + axis : NULL;
            const double *init = (axis ? y0 : x0) + (size_t)b * 3;
            double x[3] = {init[0], init[1], init[2]};
            const int push = axis == 1 && impulse != NULL;
            const int ftick = push && force_tick ? force_tick[b] : -1;
            double viol = 0.0;
            int notconv = 0, total_iters = 0;
            if (n >= 2) {
                const double zmp = x[0] - hg * x[2];
                const double below = plo[0] - zmp, above = zmp - phi[0];
                viol = fmax(fmax(below, above), 0.0);
                if (hist) { hist[0] = x[0]; hist[1] = x[1]; hist[2] = x[2]; }
                if (pf) pf[0] = (uint8_t)((below > p->viol_tol ? 4 : 0) | (above > p->viol_tol ? 8 : 0));
            }
            for (int i = 0; i + 1 < n; ++i) {
                for (int j = 0; j < N; ++j) {          /* window rows i+1 .. i+N, last row repeated (:81-88) */
                    int s = i + 1 + j;
                    s = s > n - 1 ? n - 1 : s;
                    lo[j] = plo[(size_t)s * 2];
                    hi[j] = phi[(size_t)s * 2];
                }
                double jerk;
                unsigned f = 0;
                if (mode == 0) {
                    double acc = 0.0;
                    for (int j = 0; j < N; ++j) acc += p->gain[j] * ((hi[j] + lo[j]) / 2);
                    jerk = acc + p->gain[N] * x[0] + p->gain[N + 1] * x[1] + p->gain[N + 2] * x[2];
                } else if (mode == 2) {
                    /* active-set guess: the previous tick's set moved up by one sample, new last row free */
                    /* guess = prediction of this tick's set from the last two (act = A_{i-1}, older = A_{i-2}, each in
                     * its own window coordinates) */
                    if (i == 0) { memset(act, 0, N); memset(older, 0, N); }
                    else {
                        predict_active_set(N, older, act, guess);
                        memcpy(older, act, N);
                        memcpy(act, guess, N);
                    }
                    int ok;
                    double j0;
                    const int sw = riccati_pdas_solve(p, x, lo, hi, act, gain, zplan, &j0, 64, &ok);
                    total_iters += sw;
                    if (g_sweep_log) g_sweep_log[(size_t)col * n_max + i] = (uint8_t)sw;
                    if (!ok) { ++notconv; f |= 16; }
                    f |= (act[0] < 0 ? 1u : 0u) | (act[0] > 0 ? 2u : 0u);
                    const double pred = x[0] + T * x[1] + px02 * x[2];
                    const double z0 = act[0] ? zplan[0] : clip(pred + pu00 * j0, lo[0], hi[0]);
                    jerk = (z0 - pred) / pu00;
                } else {
                    if (i == 0) {
                        for (int j = 0; j < N; ++j) tv[j] = 0.5 * (lo[j] + hi[j]);
                    } else {                              /* warm start: shift by one sample */
                        for (int j = 0; j + 1 < N; ++j) tv[j] = tv[j + 1];
                        tv[N - 1] = clip(tv[N - 1], lo[N - 1], hi[N - 1]);
                    }
                    int ok;
                    total_iters += admm_solve(p, x, lo, hi, tv, rhs, q, &ok);
                    if (!ok) { ++notconv; f |= 16; }
                    const double z0 = clip(tv[0], lo[0], hi[0]);
                    f |= (z0 == lo[0] ? 1u : 0u) | (z0 == hi[0] ? 2u : 0u);
                    const double pred = x[0] + T * x[1] + px02 * x[2];
                    jerk = (z0 - pred) / pu00;
                }
                double n0 = x[0] + T * x[1] + c2 * x[2] + c3 * jerk;     /* A x + B u0 (:199) */
                double n1 = x[1] + T * x[2] + c2 * jerk;
                double n2 = x[2] + T * jerk;
                const double zmp = n0 - hg * n2;
                const double below = lo[0] - zmp, above = zmp - hi[0];
                viol = fmax(viol, fmax(fmax(below, above), 0.0));
                f |= (below > p->viol_tol ? 4u : 0u) | (above > p->viol_tol ? 8u : 0u);
                if (i == ftick) n1 -= impulse[b];                          /* push on the y velocity (:105-106) */
                x[0] = n0; x[1] = n1; x[2] = n2;
                if (hist) { hist[(size_t)(i + 1) * 3] = n0; hist[(size_t)(i + 1) * 3 + 1] = n1; hist[(size_t)(i + 1) * 3 + 2] = n2; }
                if (pf) pf[(size_t)(i + 1) * 2] = (uint8_t)f;
            }
            if (max_violation) max_violation[b * 2 + axis] = viol;
            if (status) status[b * 2 + axis] = notconv;
            if (iters) iters[b * 2 + axis] = total_iters;
        }
        free(tv); free(act);
    }
    return 0;
}

int lipm_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
