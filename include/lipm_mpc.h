/*
 * lipm_mpc.h -- C ABI of the batched LIPM / ZMP-preview MPC engine (B200, sm_100a).
 *
 * Drop-in boundary for the hot path of
 *   plevankiem/model-predictive-control-for-bipedal-locomotion
 * The reference has no FFI: the path is plain Python methods of ZMPController
 * (src/mpc_bipedal/controllers/zmp_controller.py).  Each entry point below cites the
 * reference code it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - all floating-point data is IEEE float64 (the reference is NumPy float64 throughout);
 *   - "dev" pointers are device pointers on the plan's device, owned by the caller;
 *   - every function returns LIPM_OK (0) or a negative LIPM_E_* code and never throws;
 *     lipm_last_error() returns a thread-local message for the last failure;
 *   - calls that take a stream are asynchronous on that stream (pass 0 for the legacy stream);
 *   - a plan is immutable after creation and RE-ENTRANT: rollouts / ticks on one plan may overlap on
 *     different streams or threads as long as their output buffers are distinct (every launch carries
 *     its own work counter; only lipm_plan_stats, "the last launch", is then ambiguous);
 *   - one axis-solve of one tick of one scenario is one "QP solve"; a scenario with n bound
 *     samples costs 2*(n-1) solves (reference loop zmp_controller.py:93-104).
 */
#ifndef LIPM_MPC_H
#define LIPM_MPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lipm_plan lipm_plan_t;

/* return codes */
#define LIPM_OK 0
#define LIPM_E_INVALID (-1)     /* bad argument (message says which)                        */
#define LIPM_E_CUDA (-2)        /* a CUDA runtime call failed                               */
#define LIPM_E_UNSUPPORTED (-3) /* valid request this build cannot serve (e.g. horizon)     */
#define LIPM_E_NUMERIC (-4)     /* factorisation broke down (non-finite / non-SPD matrix)   */

/* mode of a solve / rollout -- MPCConfig.strict (config.py:39) */
#define LIPM_MODE_UNCONSTRAINED 0 /* closed form, zmp_controller.py:196-198                 */
#define LIPM_MODE_STRICT 1        /* box-constrained QP, zmp_controller.py:173-195          */

/* arithmetic of the strict mode -- north_star: "FP64 DMMA for the exactness path, and a
 * reduced-precision path with FP32 accumulate for throughput" */
#define LIPM_PRECISION_FP64 0   /* FP64 tensor pipe (DMMA); trajectories <= 1e-9 relative          */
#define LIPM_PRECISION_F16X3 1  /* tcgen05 FP16 operands split hi+lo (3 products), FP32 accumulate,
                                   FP32 ADMM state, FP64 walk state; trajectories <= 1e-4 m; N <= 152 */

/* per-sample flag bits written by rollouts: flags[b][t][axis] */
#define LIPM_FLAG_ACTIVE_LO 1u     /* applied ZMP sample sits on its lower bound (strict)   */
#define LIPM_FLAG_ACTIVE_HI 2u     /* ... on its upper bound                                */
#define LIPM_FLAG_VIOL_LO 4u       /* realised ZMP  <  z_min - viol_tol                     */
#define LIPM_FLAG_VIOL_HI 8u       /* realised ZMP  >  z_max + viol_tol                     */
#define LIPM_FLAG_NOT_CONVERGED 16u /* the QP that produced this sample hit max_iter        */

/* matrices a test / host mirror can read back with lipm_plan_export */
#define LIPM_EXPORT_PX 0        /* [N,3]   zmp_controller.py:163-165                        */
#define LIPM_EXPORT_PU 1        /* [N,N]   zmp_controller.py:166-171                        */
#define LIPM_EXPORT_ABC 2       /* [15]    A (9, row-major), B (3), C (3)  zmp_controller.py:18-20 */
#define LIPM_EXPORT_ADMM_M 3    /* [N,N]   (Q I + R Pu^-T Pu^-1 + rho I)^-1                 */
#define LIPM_EXPORT_HINV 4      /* [N,N]   (Q I + R Pu^-T Pu^-1)^-1                         */
#define LIPM_EXPORT_GAIN 5      /* [N+3]   first-jerk gain row: kz (N) then kx (3)          */
#define LIPM_EXPORT_GPX 6       /* [N,3]   R Pu^-T Pu^-1 Px                                 */

typedef struct lipm_plan_desc {
    int32_t horizon;   /* N  = MPCConfig.horizon                                            */
    double h, g, dt;   /* CoM height, gravity, sample time (dt = 1.5/N unless overridden)   */
    double Q, R;       /* tracking / jerk weights (zmp_controller.py:186, :174)             */
    double rho;        /* ADMM penalty (<=0 -> 2.0 for N <= 160, sqrt(Q (Q + lambda_max(G))) beyond) */
    double alpha;      /* ADMM over-relaxation  (<=0 -> default 1.75)                      */
    double eps;        /* ADMM stop: primal and rho*dual residual <= eps * max(1, |q|_inf)  (<=0 -> 1e-12;
                          F16X3: <=0 -> 1e-5, floor 5e-7)                                   */
    int32_t max_iter;  /* ADMM iteration cap per QP (<=0 -> 2000; F16X3: 500)                */
    double viol_tol;   /* box-violation tolerance for LIPM_FLAG_VIOL_* (<0 -> 1e-9)         */
    int32_t device;    /* CUDA device ordinal                                               */
    int32_t precision; /* LIPM_PRECISION_* of the strict mode (eps default 1e-5 for F16X3)          */
} lipm_plan_desc_t;

/*
 * Build, ON THE DEVICE, everything that is fixed for one (N, h, g, dt, Q, R):
 * A, B, C (zmp_controller.py:15-21), Px and Pu (:162-171 -- the reference rebuilds these with
 * a Python double loop on every solve), the ZMP-coordinate Hessian H = Q I + R Pu^-T Pu^-1,
 * its Cholesky-based inverses (H^-1 for the unconstrained mode, (H + rho I)^-1 for ADMM) and
 * the first-jerk gain row.  Replaces the per-call matrix construction in predict_wieber_axis.
 */
int lipm_plan_create(lipm_plan_t** out, const lipm_plan_desc_t* desc);
void lipm_plan_destroy(lipm_plan_t* plan);

/* Copy one of the plan's matrices to host memory (blocking).  `capacity` counts doubles. */
int lipm_plan_export(const lipm_plan_t* plan, int which, double* host_out, size_t capacity);

/*
 * Whole receding-horizon walks for B independent scenarios: replaces
 * ZMPController.generate_com_trajectory_wieber (zmp_controller.py:59-108) and
 * generate_state_trajectory_wieber (:110-147; pass force_tick = NULL).
 *
 *   B            scenarios
 *   n_max        allocated samples per scenario in the [.., n_max, ..] arrays
 *   n_steps      dev int32 [B]   bound samples of each scenario (2 <= n <= n_max); NULL -> n_max
 *   x0, y0       dev [B,3]       initial axis states [pos, vel, acc]
 *   z_min,z_max  dev [Bz,n_max,2] ZMP box per sample (x,y interleaved as the reference's (n,2)
 *                                arrays); bounds_batch = Bz is 1 (shared by all scenarios) or B.
 *                                The last row is repeated `horizon` times past the end (:81-88).
 *   impulse      dev [B]         dt*F_ext/m, subtracted from the y velocity after the update of
 *                                tick force_tick[b] (:105-106); NULL -> no push
 *   force_tick   dev int32 [B]   the reference uses n_steps//2 (:90); negative -> no push
 * outputs (any may be NULL to skip)
 *   x_hist,y_hist dev [B,n_max,3] state histories; sample 0 is the initial state
 *   flags        dev uint8 [B,n_max,2]
 *   max_violation dev [B,2]      per axis: max over samples of max(z_min - zmp, zmp - z_max, 0)
 *   status       dev int32 [B,2] per axis: 1 if a QP hit max_iter (0 = all converged).  Like the reference,
 *                                which raises at the first QP without a solution (:193-194), the walk of
 *                                that axis stops at that sample (flagged NOT_CONVERGED); later samples
 *                                are left untouched
 *   iters        dev int32 [B,2] per axis: ADMM iterations executed, summed over ticks
 */
int lipm_rollout(lipm_plan_t* plan, int mode, int B, int n_max, const int32_t* n_steps,
                 const double* x0, const double* y0, const double* z_min, const double* z_max,
                 int bounds_batch, const double* impulse, const int32_t* force_tick,
                 double* x_hist, double* y_hist, uint8_t* flags, double* max_violation,
                 int32_t* status, int32_t* iters, void* stream);

/*
 * Same as lipm_rollout, with a TABLE of bound series and one index per scenario: z_min / z_max are
 * dev [V, n_max, 2] (V = bounds_batch series) and bounds_index is dev int32 [B] with values in [0, V).
 * Parameter sweeps use it to run scenarios with different footstep plans (step length, foot spread --
 * generators/footstep_generator.py:19-49) in one launch; n_steps gives each scenario its series' length.
 */
int lipm_rollout_indexed(lipm_plan_t* plan, int mode, int B, int n_max, const int32_t* n_steps,
                         const double* x0, const double* y0, const double* z_min, const double* z_max,
                         int bounds_batch, const int32_t* bounds_index, const double* impulse,
                         const int32_t* force_tick, double* x_hist, double* y_hist, uint8_t* flags,
                         double* max_violation, int32_t* status, int32_t* iters, void* stream);

/*
 * Per-scenario outcome of a walk: what a sweep all-gathers across GPUs (32 bytes).  The reference computes no
 * outcome; its drivers plot  C . state  against the bounds (scripts/run_compare_resistance.py:173-216).
 */
typedef struct lipm_outcome {
    double max_violation;  /* max over axes and samples of max(z_min - zmp, zmp - z_max, 0)              */
    double lean_max;       /* max over axes and samples of |CoM - ZMP| = (h/g) |CoM acceleration|        */
    int32_t status;        /* QPs that hit max_iter, both axes (0 = every QP converged)                   */
    int32_t fell;          /* lean_max > fall_lean (a definition of this library, see lipm_rollout_io_t) */
    int64_t iters;         /* ADMM iterations, both axes                                                  */
} lipm_outcome_t;

/*
 * Arguments of lipm_rollout_ex: lipm_rollout_indexed plus the read-outs the reference's drivers compute on the
 * host after the walk -- the CoM path (zmp_controller.py:108 `com`), the realised ZMP  C . state
 * (scripts/run_mpc.py:294, run_compare_resistance.py:173-174) -- and the packed per-scenario outcome.  They are
 * written by the walk kernels themselves (the ZMP is in registers there), not by a second pass over the histories.
 * Zero-initialise the struct; any output may be NULL.
 */
typedef struct lipm_rollout_io {
    int32_t B, n_max;
    const int32_t* n_steps;        /* dev [B] or NULL                                                    */
    const double *x0, *y0;         /* dev [B,3]                                                          */
    const double *z_min, *z_max;   /* dev [bounds_batch, n_max, 2]                                       */
    int32_t bounds_batch;          /* 1, B, or the number of series V when bounds_index is given         */
    const int32_t* bounds_index;   /* dev [B] or NULL                                                    */
    const double* impulse;         /* dev [B] or NULL                                                    */
    const int32_t* force_tick;     /* dev [B] or NULL                                                    */
    /* outputs */
    double *x_hist, *y_hist;       /* dev [B,n_max,3]                                                    */
    double* com;                   /* dev [B,n_max,2]  (x_hist[..,0], y_hist[..,0])                      */
    double* zmp;                   /* dev [B,n_max,2]  C . state per axis                                */
    uint8_t* flags;                /* dev [B,n_max,2]                                                    */
    double* max_violation;         /* dev [B,2]                                                          */
    double* lean_max;              /* dev [B,2] per axis max |CoM - ZMP|                                 */
    int32_t* status;               /* dev [B,2]                                                          */
    int32_t* iters;                /* dev [B,2]                                                          */
    lipm_outcome_t* outcome;       /* dev [B] packed record (both axes combined)                         */
    double fall_lean;              /* fell = lean_max > fall_lean; <= 0 -> h tan(30 deg): the pendulum leans more
                                      than 30 degrees.  In strict mode the ZMP cannot leave the sole, so an
                                      unrecoverable push shows up as the CoM running away from it           */
} lipm_rollout_io_t;

int lipm_rollout_ex(lipm_plan_t* plan, int mode, const lipm_rollout_io_t* io, void* stream);

/*
 * Process- or plan-level switches that are off by default.
 *   LIPM_OPT_KEEP_POOL_MEMORY (plan may be NULL; value = device ordinal): raise the release threshold of that
 *       device's default stream-ordered memory pool so that the scratch buffers of the rollouts stay cached
 *       between calls (with the default threshold the pool returns its memory to the driver at every
 *       synchronisation; measured: 0.1 - 1 s on one call in three).  This changes the memory behaviour of the
 *       whole process on that device, so the library never does it on its own.
 *   LIPM_OPT_TICK_RANGES (plan required; value = 0 automatic, 1..8 forced): how many consecutive tick ranges the
 *       FP64 strict kernel cuts a walk into for load balance.  Results do not depend on it (tested).
 *   LIPM_OPT_SINGLE_CTA_F16 (plan required; value 0 / 1): run the reduced-precision walks (N <= 152) on the single-CTA
 *       tcgen05 kernel instead of the CTA-pair one (cta_group::2).  Same arithmetic, same bits (tested); kept for
 *       comparison and as the reference point of the pair kernel's speed-up.
 */
#define LIPM_OPT_KEEP_POOL_MEMORY 1
#define LIPM_OPT_TICK_RANGES 2
#define LIPM_OPT_SINGLE_CTA_F16 3
#define LIPM_OPT_STRICT_SOLVER 4   /* plan required; LIPM_SOLVER_* of the FP64 strict mode                        */
#define LIPM_OPT_RICCATI_LANES 5   /* plan required; columns per warp of the active-set kernel (0 automatic,
                                      else a power of two <= 32); results do not depend on it                    */
#define LIPM_OPT_RICCATI_SCRATCH 7  /* plan required; where the active-set kernel keeps the feed-forward terms of its
                                      table stages: 0 automatic, 1 shared memory, 2 global scratch; results do not
                                      depend on it                                                               */
#define LIPM_OPT_FEW_PIN_ROWS 8     /* plan required; active-set kernel: sweeps with at most this many pinned rows solve
                                      the unconstrained problem with the per-plan tables and correct it with the dense
                                      inverse Hessian (0 = never, the default: in the one-thread-per-column kernel the
                                      Riccati recursion is faster; up to 16); both give the same optimum           */
#define LIPM_OPT_ROW_PARALLEL 9     /* plan required; 1 (default): FP64 active-set walks on shared or indexed bound series of up
                                      to 16384 columns start on the row-parallel kernel (csrc/lipm_rollout_rowpar.cu); 2: whatever
                                      the batch size; 0: never (the whole walk on the thread-per-column Riccati kernel).  Same
                                      optimum either way (walks agree to 1e-9, flags bit for bit: tested)                    */
#define LIPM_OPT_WALK_CHUNKS 10     /* plan required; FP64 active-set walks that start on the row-parallel kernel: the batch is cut into
                                      this many chunks of the processing order whose later stages overlap on side streams of the
                                      library (0 automatic: 4 from 4096 scenarios on when the call carries a push, 1: everything on the caller's stream, up to 4);
                                      every column does the same arithmetic whatever the value (bit-identical: tested)         */
#define LIPM_OPT_MAX_SWEEPS 6      /* plan required; active-set corrections per QP before it is reported as
                                      not converged (0 -> 64)                                                    */

/* solver of the FP64 strict mode (LIPM_OPT_STRICT_SOLVER).  Both return the unique minimiser of the reference's
 * QP (zmp_controller.py:173-191); they share no arithmetic, so their agreement is a test (tests/test_cuda_parity.py). */
#define LIPM_SOLVER_ACTIVE_SET 0   /* default: primal-dual active set, sub-problems solved exactly by an O(N) Riccati
                                      sweep over the 3-state model (csrc/lipm_rollout_riccati.cu); `iters` counts
                                      sweep pairs                                                                */
#define LIPM_SOLVER_ADMM 1         /* fixed-rho ADMM on the condensed QP, FP64 tensor pipe (csrc/lipm_rollout_strict.cu);
                                      `iters` counts ADMM iterations; rho / alpha / eps / max_iter apply           */
int lipm_set_option(lipm_plan_t* plan, int option, long long value);

/*
 * One control tick for a batch of independent axis problems: replaces
 * ZMPController.predict_wieber_axis (zmp_controller.py:149-201).
 *
 *   state      dev [B,3]      current axis state
 *   z_min,z_max dev [Bz,N]    bound window rows i+1 .. i+N of that axis (Bz = 1 or B)
 *   next_state dev [B,3]      A x + B u0
 *   zmp_plan   dev [B,N] or NULL  the optimal previewed ZMP sequence Px x + Pu J
 *   jerk_plan  dev [B,N] or NULL  the optimal jerk sequence J  (X_burst of the reference)
 *   status     dev int32 [B] or NULL  0 converged / 1 hit max_iter (strict mode)
 *   iters      dev int32 [B] or NULL  ADMM iterations executed
 */
int lipm_solve_tick(lipm_plan_t* plan, int mode, int B, const double* state, const double* z_min,
                    const double* z_max, int bounds_batch, double* next_state, double* zmp_plan,
                    double* jerk_plan, int32_t* status, int32_t* iters, void* stream);

/*
 * Work accounting of the last strict launch on this plan (blocking read-back):
 *   host_out[0] = ADMM iterations executed by 8-column solver groups (each occupies 8 column
 *                 slots of the tensor pipe, converged or not); host_out[1] = 8-column units walked;
 *   host_out[2], host_out[3] = debug cycle counters (zero unless LIPM_STRICT_SKIP has bit 16).
 *   host_out must hold 4 values.
 * Useful-work fraction = sum(iters) / (8 * host_out[0]).
 */
int lipm_plan_stats(const lipm_plan_t* plan, uint64_t* host_out);

/*
 * Footstep plans and ZMP bound series of V walk variants, generated on the device (one thread per variant):
 * generate_footsteps (generators/footstep_generator.py:19-49) + CoPGenerator.generate_cop_trajectory
 * (generators/cop_generator.py:73-115), bit-identical to the host generators (same floating-point clock).
 *   distance, step_length, foot_spread   dev [V]
 *   z_min, z_max   dev [V, n_max, 2];  states dev int8 [V, n_max] or NULL (0 STANDING, 1 DOUBLE, 2 SINGLE)
 *   n_out          dev int32 [V]  samples of each series; -needed when a series does not fit n_max
 * The table feeds lipm_rollout_indexed (bounds_index = variant of each scenario, n_steps = n_out[variant]).
 */
int lipm_generate_bounds(int V, const double* distance, const double* step_length, const double* foot_spread,
                         double ssp_duration, double dsp_duration, double standing_duration, double dt, int n_max,
                         double* z_min, double* z_max, int8_t* states, int32_t* n_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Herdt free-footstep MPC (method = "herdt"): ZMPController.predict_herdt_joint (zmp_controller.py:533-826)
 * for B scenarios that share a support timeline.  The rollout bookkeeping of
 * generate_com_trajectory_herdt (:435-531: swing-foot interpolation, side toggle, footstep commit, push)
 * stays with the caller, tick by tick, on device tensors.
 * ------------------------------------------------------------------------------------------------ */
typedef struct herdt_plan herdt_plan_t;
typedef struct herdt_timeline herdt_timeline_t;

typedef struct herdt_plan_desc {
    int32_t horizon;                /* N (<= 152: N + footstep slots must fit the resident-matrix kernel) */
    double h, g, dt;
    double alpha, beta, gamma;      /* jerk / velocity / ZMP weights (config.py:42-44, zmp_controller.py:605-607) */
    double foot_length, foot_width; /* ZMP box around the supporting foot (:661-680)                      */
    double foot_spread;             /* standing hull (:715-750)                                           */
    double rho, relax, eps;         /* ADMM penalty (<=0 -> 6), over-relaxation (<=0 -> 1.8), stop (<=0 -> 1e-12) */
    int32_t max_iter;               /* <=0 -> 600 (a healthy tick needs 60..150 from a cold start)         */
    int32_t device;
    int32_t stall_window;           /* a pair whose residual lost < 10 % over this many iterations stops (<=0 -> 32;
                                       >= max_iter: never)                                                 */
    double accept_eps;              /* a pair that stops WITHOUT reaching eps is still applied (status 2) when its residuals pass
                                       this looser relative tolerance: 0 -> 1e-5 = the tolerance at which the reference's solver
                                       (OSQP through cvxpy) declares "solved" and the reference applies the point
                                       (zmp_controller.py:792-805); < 0 -> never (every such pair falls back, :796-804) */
} herdt_plan_desc_t;

/* Per (N, h, g, dt, alpha, beta): T = Pzu^-1, V = Pvu Pzu^-1, W = alpha T'T + beta V'V, on the device. */
int herdt_plan_create(herdt_plan_t** out, const herdt_plan_desc_t* desc);
void herdt_plan_destroy(herdt_plan_t* plan);
int herdt_plan_export_w(const herdt_plan_t* plan, double* host_out, size_t capacity);

/*
 * Per support timeline: for every tick the layout of the preview window (host arrays, :581-603, :682-698):
 *   phase_of_row [n_ticks, N] int8   -1: row supported by the current foot (U_c), k >= 0: by future footstep k (U)
 *   row_kind     [n_ticks, N] int8    0: unconstrained, 1: ZMP box around its foot, 2: standing hull of both feet
 *   m_per_tick   [n_ticks]   int32   future footsteps in the window (m_max = their maximum)
 * Builds the Hessian of every tick, its shifted inverse and the DMMA fragment image, on the device.
 */
int herdt_timeline_create(herdt_plan_t* plan, int n_ticks, int m_max, const int8_t* phase_of_row,
                          const int8_t* row_kind, const int32_t* m_per_tick, herdt_timeline_t** out);
void herdt_timeline_destroy(herdt_timeline_t* timeline);
int herdt_timeline_size(const herdt_timeline_t* timeline);   /* padded variable count per axis */

/*
 * One control tick (predict_herdt_joint) for B scenarios of the timeline:
 *   x, y        dev [B,3]   axis states;   fc dev [B,2] current support foot;
 *   air         dev [B,2]   swing-foot interpolation (x_airc, y_airc) or NULL: a scenario whose QP does not
 *               converge falls back like the reference (:796-804): zero jerk, footstep := air
 *   spread      dev [B]     per-scenario foot_spread of the standing hull, or NULL (-> plan value)
 *   vref        dev [vref_batch, n_ext, 2] reference velocities, already padded by the caller (:465-467);
 *               window row j reads sample tick + 1 + j
 *   foot_side   0 = left, 1 = right (selects the standing hull; the polygon is passed explicitly)
 *   poly_verts  HOST [n_verts, 2] reachable polygon of the next footstep relative to fc, counter-clockwise
 *               (config.py:57-82 through ConvexHull, zmp_controller.py:773-790)
 *   warm        0: cold start (as the reference: a fresh OSQP problem per tick);  1: start from the previous
 *               tick's ADMM iterate kept in the timeline (only honoured when the previous call was tick - 1
 *               with the same B);  2: same, and the previous tick's first future footstep has just become
 *               the current phase (footstep variables move down by one)
 * outputs: next_x, next_y dev [B,3]; f1 dev [B,2] first free footstep (NaN when the window holds none);
 *          solution dev [2B, size] or NULL ((e, F) per axis, x and y columns of a scenario adjacent);
 *          status, iters dev int32 [2B] or NULL; status: 0 converged to eps, 2 applied at accept_eps, 1 failed (fallback).
 */
int herdt_tick(herdt_timeline_t* timeline, int tick, int B, const double* x, const double* y, const double* fc,
               const double* air, const double* spread, const double* vref, int vref_batch, int n_ext, int foot_side,
               const double* poly_verts, int n_verts, int warm, double* next_x, double* next_y, double* f1,
               double* solution, int32_t* status, int32_t* iters, void* stream);

/* A whole Herdt walk for B scenarios that share the timeline: the tick loop of generate_com_trajectory_herdt
 * (zmp_controller.py:475-525) enqueued on `stream` without returning to the host between ticks -- per tick
 * herdt_tick, then the swing-foot interpolation (:495-498), the footstep commit at the end of a single support
 * (:501-517), the push (:522-523), the histories and the running totals.
 *   n           samples of the walk (n - 1 ticks, n - 1 <= n_ticks of the timeline)
 *   x0, y0      dev [B,3];  fc0 dev [B,2] initial supporting foot;  spread dev [B] or NULL
 *   vref        dev [vref_batch, n_ext, 2] as for herdt_tick
 *   foot_side, steps_left, commit, new_run   HOST arrays [n-1]: supporting side of each tick (0 left, 1 right),
 *               ticks until the next state change (find_nb_steps(...)[i][0]), "this tick ends a single
 *               support" and "a new support run starts at this tick" (warm start bookkeeping)
 *   poly_left / poly_right   HOST [n_left|n_right, 2] reachable polygons of the next footstep per supporting side
 *   warm_start  0: every tick starts cold (as the reference);  1: from the previous tick's iterate
 *   impulse     dev [B] = dt * F_ext / m applied to the y velocity after tick `force_tick`, or NULL
 * outputs: x_hist, y_hist dev [B,n,3]; foot_hist dev [B,n,2]; status dev int32 [B]: low 16 bits = ticks whose joint QP
 *          failed (they fall back like the reference, :796-804), high 16 bits = ticks applied at accept_eps;
 *          iters dev int64 [B]; both may be NULL.
 * Walks of DIFFERENT timelines may be enqueued on different streams and run concurrently. */
int herdt_rollout(herdt_timeline_t* timeline, int B, int n, const double* x0, const double* y0, const double* fc0,
                  const double* spread, const double* vref, int vref_batch, int n_ext, const int8_t* foot_side,
                  const int32_t* steps_left, const uint8_t* commit, const uint8_t* new_run, const double* poly_left,
                  int n_left, const double* poly_right, int n_right, int warm_start, const double* impulse, int force_tick,
                  double* x_hist, double* y_hist, double* foot_hist, int32_t* status, long long* iters, void* stream);

/* Library / build information. */
const char* lipm_last_error(void);
const char* lipm_version(void);
/* Largest horizon of the reduced-precision strict kernel (LIPM_PRECISION_F16X3). */
int lipm_max_reduced_horizon(void);
/* Largest horizon the strict kernels of this build serve (the matrix is resident in shared
 * memory up to N = 160 and streamed from L2 beyond that). */
int lipm_max_strict_horizon(void);

#ifdef __cplusplus
}
#endif
#endif /* LIPM_MPC_H */
