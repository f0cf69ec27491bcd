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

/*
 * Walks of B scenarios.  Same argument meaning as lipm_rollout in include/lipm_mpc.h, with host
 * pointers.  mode 0 = non-strict (zmp_controller.py:196-199), 1 = strict.  Returns 0.
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
        double *tv = (double *)malloc(sizeof(double) * N * 5);
        double *lo = tv + N, *hi = tv + 2 * N, *rhs = tv + 3 * N, *q = tv + 4 * N;
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
        free(tv);
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
