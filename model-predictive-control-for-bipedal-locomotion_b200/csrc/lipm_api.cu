// C ABI entry points (include/lipm_mpc.h).  Argument checking, plan lifetime, dispatch.
#include "lipm_internal.cuh"

#include <cmath>
#include <cstring>
#include <new>

namespace lipm {

namespace {
thread_local char g_error[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

struct DeviceGuard {
    int previous = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&previous) != cudaSuccess) return;
        ok = (previous == device) || (cudaSetDevice(device) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (previous >= 0) cudaSetDevice(previous);
    }
};

void free_plan_buffers(lipm_plan* p) {
    void* ptrs[] = {p->d_px, p->d_pu, p->d_abc, p->d_tinv, p->d_g, p->d_gpx, p->d_work,
                    p->d_hinv, p->d_minv, p->d_mfrag, p->d_hfrag, p->d_gain, p->d_info, p->d_stats, p->d_bf16, p->d_f16s, p->d_ric_tab, p->d_ric_aux, p->d_kx};
    for (void* q : ptrs)
        if (q) cudaFree(q);
}

// One thread per scenario: the per-axis results of a walk -> the packed record a sweep all-gathers.
__global__ void k_pack_outcomes(int B, const double* __restrict__ viol, const double* __restrict__ lean,
                                const int32_t* __restrict__ status, const int32_t* __restrict__ iters, double fall_lean,
                                lipm_outcome_t* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double2 v = reinterpret_cast<const double2*>(viol)[b];
    const double2 l = reinterpret_cast<const double2*>(lean)[b];
    const int2 s = reinterpret_cast<const int2*>(status)[b];
    const int2 it = reinterpret_cast<const int2*>(iters)[b];
    lipm_outcome_t r;
    r.max_violation = fmax(v.x, v.y);
    r.lean_max = fmax(l.x, l.y);
    r.status = s.x + s.y;
    r.fell = r.lean_max > fall_lean ? 1 : 0;
    r.iters = (int64_t)it.x + (int64_t)it.y;
    out[b] = r;
}

}  // namespace

int launch_pack_outcomes(int B, const double* max_violation, const double* lean_max, const int32_t* status, const int32_t* iters,
                         double fall_lean, lipm_outcome_t* outcome, cudaStream_t stream) {
    if (B <= 0) return LIPM_OK;
    k_pack_outcomes<<<(B + 255) / 256, 256, 0, stream>>>(B, max_violation, lean_max, status, iters, fall_lean, outcome);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}
}  // namespace lipm

using namespace lipm;

extern "C" {

const char* lipm_last_error(void) { return g_error; }

const char* lipm_version(void) { return "lipm_mpc 0.4 (sm_100a; fp64 active-set Riccati, fp64 DMMA ADMM, tcgen05 f16x3 ADMM, gain-row unconstrained)"; }

int lipm_max_reduced_horizon(void) { return kMaxReducedStreamHorizon; }

int lipm_max_strict_horizon(void) { return kMaxStrictHorizon; }

int lipm_plan_create(lipm_plan_t** out, const lipm_plan_desc_t* desc) {
    if (!out || !desc) { set_error("lipm_plan_create: null argument"); return LIPM_E_INVALID; }
    *out = nullptr;
    lipm_plan_desc_t d = *desc;
    if (d.horizon < 2 || d.horizon > 4096) { set_error("horizon %d out of range [2, 4096]", d.horizon); return LIPM_E_INVALID; }
    if (!(d.h > 0) || !(d.g > 0) || !(d.dt > 0) || !(d.Q > 0) || !(d.R > 0) ||
        !std::isfinite(d.h) || !std::isfinite(d.g) || !std::isfinite(d.dt) || !std::isfinite(d.Q) || !std::isfinite(d.R)) {
        set_error("h, g, dt, Q, R must be finite and positive (got %g %g %g %g %g)", d.h, d.g, d.dt, d.Q, d.R);
        return LIPM_E_INVALID;
    }
    if (!(d.rho > 0)) d.rho = d.horizon <= kMaxResidentHorizon ? 2.0 : 0.0;   // 0: from the spectrum (lipm_plan.cu)
    if (!(d.alpha > 0)) d.alpha = 1.75;   // tools/admm_tune.py: 26.8 iterations per solve on the push sweep (1.7: 27.5, 1.6: 29.5), neutral or better on cold ticks
    if (d.alpha >= 2.0) { set_error("alpha must be in (0, 2), got %g", d.alpha); return LIPM_E_INVALID; }
    if (d.precision != LIPM_PRECISION_FP64 && d.precision != LIPM_PRECISION_F16X3) {
        set_error("unknown precision %d", d.precision);
        return LIPM_E_INVALID;
    }
    if (d.precision == LIPM_PRECISION_F16X3) {
        if (d.horizon > kMaxReducedStreamHorizon) {
            set_error("precision F16X3: horizon %d exceeds the reduced-precision kernels' limit %d", d.horizon, kMaxReducedStreamHorizon);
            return LIPM_E_UNSUPPORTED;
        }
        if (!(d.eps > 0)) d.eps = 1e-5;
        if (d.max_iter <= 0) d.max_iter = 500;   // ~8 iterations per solve on average; a column that needs more has diverged
        if (d.eps < 5e-7) d.eps = 5e-7;     // FP32 iterates: below this the stop test sits in the rounding noise
    }
    if (!(d.eps > 0)) d.eps = 1e-12;
    if (d.max_iter <= 0) d.max_iter = 2000;
    if (d.viol_tol < 0 || !std::isfinite(d.viol_tol)) d.viol_tol = 1e-9;
    const double pu00 = std::pow(d.dt, 3) / 6.0 - d.dt * d.h / d.g;
    if (pu00 == 0.0) { set_error("Pu[0,0] = dt^3/6 - dt*h/g is zero: Pu is singular"); return LIPM_E_INVALID; }

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_error("no CUDA device available (this library has no CPU path)");
        return LIPM_E_CUDA;
    }
    if (d.device < 0 || d.device >= count) { set_error("device %d out of range (%d devices)", d.device, count); return LIPM_E_INVALID; }
    DeviceGuard guard(d.device);
    if (!guard.ok) { set_error("cannot select device %d", d.device); return LIPM_E_CUDA; }

    lipm_plan* plan = new (std::nothrow) lipm_plan();
    if (!plan) { set_error("out of host memory"); return LIPM_E_INVALID; }
    std::memset(plan, 0, sizeof(*plan));
    plan->desc = d;
    plan->N = d.horizon;
    plan->NP = padded8(d.horizon);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d.device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        delete plan;
        return LIPM_E_CUDA;
    }
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", d.device, prop.major, prop.minor);
        delete plan;
        return LIPM_E_UNSUPPORTED;
    }
    plan->sm_count = prop.multiProcessorCount;
    const int rc = build_plan_matrices(plan);
    if (rc != LIPM_OK) {
        free_plan_buffers(plan);
        delete plan;
        return rc;
    }
    *out = plan;
    return LIPM_OK;
}

void lipm_plan_destroy(lipm_plan_t* plan) {
    if (!plan) return;
    DeviceGuard guard(plan->desc.device);
    free_plan_buffers(plan);
    delete plan;
}

int lipm_plan_export(const lipm_plan_t* plan, int which, double* host_out, size_t capacity) {
    if (!plan || !host_out) { set_error("lipm_plan_export: null argument"); return LIPM_E_INVALID; }
    const size_t N = (size_t)plan->N;
    const double* src = nullptr;
    size_t count = 0;
    switch (which) {
        case LIPM_EXPORT_PX: src = plan->d_px; count = N * 3; break;
        case LIPM_EXPORT_PU: src = plan->d_pu; count = N * N; break;
        case LIPM_EXPORT_ABC: src = plan->d_abc; count = 15; break;
        case LIPM_EXPORT_ADMM_M: src = plan->d_minv; count = N * N; break;
        case LIPM_EXPORT_HINV: src = plan->d_hinv; count = N * N; break;
        case LIPM_EXPORT_GAIN: src = plan->d_gain; count = N + 3; break;
        case LIPM_EXPORT_GPX: src = plan->d_gpx; count = N * 3; break;
        default: set_error("lipm_plan_export: unknown selector %d", which); return LIPM_E_INVALID;
    }
    if (capacity < count) { set_error("lipm_plan_export: need %zu doubles, got %zu", count, capacity); return LIPM_E_INVALID; }
    DeviceGuard guard(plan->desc.device);
    LIPM_CUDA_TRY(cudaMemcpy(host_out, src, count * sizeof(double), cudaMemcpyDeviceToHost));
    return LIPM_OK;
}

int lipm_plan_stats(const lipm_plan_t* plan, uint64_t* host_out) {
    if (!plan || !host_out) { set_error("lipm_plan_stats: null argument"); return LIPM_E_INVALID; }
    DeviceGuard guard(plan->desc.device);
    LIPM_CUDA_TRY(cudaMemcpy(host_out, plan->d_stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return LIPM_OK;
}

int lipm_set_option(lipm_plan_t* plan, int option, long long value) {
    switch (option) {
        case LIPM_OPT_KEEP_POOL_MEMORY: {
            int count = 0;
            if (cudaGetDeviceCount(&count) != cudaSuccess || value < 0 || value >= count) {
                set_error("lipm_set_option(KEEP_POOL_MEMORY): bad device %lld", value);
                return LIPM_E_INVALID;
            }
            cudaMemPool_t pool = nullptr;
            LIPM_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, (int)value));
            unsigned long long keep = ~0ull;
            LIPM_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            return LIPM_OK;
        }
        case LIPM_OPT_TICK_RANGES:
            if (!plan || value < 0 || value > 8) { set_error("lipm_set_option(TICK_RANGES): need a plan and a value in [0, 8]"); return LIPM_E_INVALID; }
            plan->force_segments = (int)value;
            return LIPM_OK;
        case LIPM_OPT_SINGLE_CTA_F16:
            if (!plan || (value != 0 && value != 1)) { set_error("lipm_set_option(SINGLE_CTA_F16): need a plan and 0 / 1"); return LIPM_E_INVALID; }
            plan->single_cta_f16 = (int)value;
            return LIPM_OK;
        case LIPM_OPT_STRICT_SOLVER:
            if (!plan || (value != LIPM_SOLVER_ACTIVE_SET && value != LIPM_SOLVER_ADMM)) { set_error("lipm_set_option(STRICT_SOLVER): need a plan and LIPM_SOLVER_ACTIVE_SET / LIPM_SOLVER_ADMM"); return LIPM_E_INVALID; }
            plan->solver = (int)value;
            return LIPM_OK;
        case LIPM_OPT_RICCATI_LANES:
            if (!plan || value < 0 || value > 32 || (value & (value - 1)) != 0) { set_error("lipm_set_option(RICCATI_LANES): need a plan and 0 or a power of two <= 32"); return LIPM_E_INVALID; }
            plan->riccati_lanes = (int)value;
            return LIPM_OK;
        case LIPM_OPT_ROW_PARALLEL:
            if (!plan || value < 0 || value > 2) { set_error("lipm_set_option(ROW_PARALLEL): need a plan and 0 / 1 / 2"); return LIPM_E_INVALID; }
            plan->rowpar_off = value == 0 ? 1 : (value == 2 ? -1 : 0);
            return LIPM_OK;
        case LIPM_OPT_WALK_CHUNKS:
            if (!plan || value < 0 || value > kRowparMaxChunks) { set_error("lipm_set_option(WALK_CHUNKS): need a plan and a value in [0, %d]", kRowparMaxChunks); return LIPM_E_INVALID; }
            plan->rowpar_chunks = (int)value;
            return LIPM_OK;
        case LIPM_OPT_FEW_PIN_ROWS:
            if (!plan || value < 0 || value > 16) { set_error("lipm_set_option(FEW_PIN_ROWS): need a plan and a value in [0, 16]"); return LIPM_E_INVALID; }
            plan->riccati_few_pins = (int)value;
            return LIPM_OK;
        case LIPM_OPT_RICCATI_SCRATCH:
            if (!plan || value < 0 || value > 2) { set_error("lipm_set_option(RICCATI_SCRATCH): need a plan and 0 (automatic), 1 (shared memory) or 2 (global)"); return LIPM_E_INVALID; }
            plan->riccati_scratch = (int)value;
            return LIPM_OK;
        case LIPM_OPT_MAX_SWEEPS:
            if (!plan || value < 0 || value > 100000) { set_error("lipm_set_option(MAX_SWEEPS): need a plan and a value in [0, 100000]"); return LIPM_E_INVALID; }
            plan->riccati_max_sweeps = (int)value;
            return LIPM_OK;
        default:
            set_error("lipm_set_option: unknown option %d", option);
            return LIPM_E_INVALID;
    }
}

int lipm_rollout(lipm_plan_t* plan, int mode, int B, int n_max, const int32_t* n_steps, const double* x0,
                 const double* y0, const double* z_min, const double* z_max, int bounds_batch,
                 const double* impulse, const int32_t* force_tick, double* x_hist, double* y_hist,
                 uint8_t* flags, double* max_violation, int32_t* status, int32_t* iters, void* stream) {
    return lipm_rollout_indexed(plan, mode, B, n_max, n_steps, x0, y0, z_min, z_max, bounds_batch, nullptr, impulse,
                                force_tick, x_hist, y_hist, flags, max_violation, status, iters, stream);
}

int lipm_rollout_indexed(lipm_plan_t* plan, int mode, int B, int n_max, const int32_t* n_steps, const double* x0,
                         const double* y0, const double* z_min, const double* z_max, int bounds_batch,
                         const int32_t* bounds_index, const double* impulse, const int32_t* force_tick,
                         double* x_hist, double* y_hist, uint8_t* flags, double* max_violation, int32_t* status,
                         int32_t* iters, void* stream) {
    lipm_rollout_io_t io;
    std::memset(&io, 0, sizeof(io));
    io.B = B; io.n_max = n_max; io.n_steps = n_steps; io.x0 = x0; io.y0 = y0; io.z_min = z_min; io.z_max = z_max;
    io.bounds_batch = bounds_batch; io.bounds_index = bounds_index; io.impulse = impulse; io.force_tick = force_tick;
    io.x_hist = x_hist; io.y_hist = y_hist; io.flags = flags; io.max_violation = max_violation; io.status = status;
    io.iters = iters;
    return lipm_rollout_ex(plan, mode, &io, stream);
}

int lipm_rollout_ex(lipm_plan_t* plan, int mode, const lipm_rollout_io_t* io, void* stream_) {
    if (!plan || !io) { set_error("lipm_rollout: null plan / arguments"); return LIPM_E_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int B = io->B, n_max = io->n_max, bounds_batch = io->bounds_batch;
    if (mode != LIPM_MODE_STRICT && mode != LIPM_MODE_UNCONSTRAINED) { set_error("lipm_rollout: bad mode %d", mode); return LIPM_E_INVALID; }
    if (B < 0 || n_max < 2) { set_error("lipm_rollout: need B >= 0 and n_max >= 2 (got %d, %d)", B, n_max); return LIPM_E_INVALID; }
    if (B == 0) return LIPM_OK;
    if (!io->x0 || !io->y0 || !io->z_min || !io->z_max) { set_error("lipm_rollout: x0, y0, z_min, z_max are required"); return LIPM_E_INVALID; }
    if (!io->bounds_index && bounds_batch != 1 && bounds_batch != B) { set_error("lipm_rollout: bounds_batch must be 1 or B (got %d, B=%d)", bounds_batch, B); return LIPM_E_INVALID; }
    if (io->bounds_index && bounds_batch < 1) { set_error("lipm_rollout_indexed: bounds_batch must be the number of series (got %d)", bounds_batch); return LIPM_E_INVALID; }
    if ((long long)B * 2 > 0x7fffffffLL / 8) { set_error("lipm_rollout: batch too large"); return LIPM_E_INVALID; }
    DeviceGuard guard(plan->desc.device);

    ColumnBatch cb;
    std::memset(&cb, 0, sizeof(cb));
    cb.ncols = 2 * B;
    cb.B = B;
    cb.bounds_batch = bounds_batch;
    cb.bounds_index = io->bounds_index;
    cb.bound_scn_stride = (long long)n_max * 2;
    cb.bound_stride = 2;
    cb.bound_axis_offset = 1;
    cb.bound_index_bias = 1;
    cb.n_max = n_max;
    cb.n_steps = io->n_steps;
    cb.n_uniform = n_max;
    cb.ticks_override = 0;
    cb.init[0] = io->x0; cb.init[1] = io->y0;
    cb.z_min = io->z_min; cb.z_max = io->z_max;
    cb.impulse = io->impulse; cb.force_tick = io->force_tick;
    cb.hist[0] = io->x_hist; cb.hist[1] = io->y_hist;
    cb.write_initial = 1;
    cb.hist_scn_stride = (long long)n_max * 3;
    cb.hist_index_bias = 0;
    cb.flags = io->flags;
    cb.max_violation = io->max_violation;
    cb.lean_max = io->lean_max;
    cb.status = io->status;
    cb.iters = io->iters;
    cb.com = io->com;
    cb.zmp = io->zmp;
    cb.zmp_plan = nullptr;

    // the packed record needs all four per-axis results: stage the ones the caller did not ask for
    StreamBuffer axis_scratch;
    if (io->outcome) {
        const size_t per = (size_t)B * 2;
        const size_t need = (cb.max_violation ? 0 : per * sizeof(double)) + (cb.lean_max ? 0 : per * sizeof(double)) +
                            (cb.status ? 0 : per * sizeof(int32_t)) + (cb.iters ? 0 : per * sizeof(int32_t));
        if (need) {
            LIPM_CUDA_TRY(axis_scratch.alloc(need, stream));
            unsigned char* p = axis_scratch.as<unsigned char>();
            if (!cb.max_violation) { cb.max_violation = reinterpret_cast<double*>(p); p += per * sizeof(double); }
            if (!cb.lean_max) { cb.lean_max = reinterpret_cast<double*>(p); p += per * sizeof(double); }
            if (!cb.status) { cb.status = reinterpret_cast<int32_t*>(p); p += per * sizeof(int32_t); }
            if (!cb.iters) { cb.iters = reinterpret_cast<int32_t*>(p); p += per * sizeof(int32_t); }
        }
    }

    int rc;
    if (mode != LIPM_MODE_STRICT) {
        rc = launch_unconstrained_rollout(plan, cb, stream);
    } else {
        // strict walks: group scenarios that will need similar iteration counts (lipm_order.cu)
        const bool active_set = plan->desc.precision != LIPM_PRECISION_F16X3 && plan->solver != LIPM_SOLVER_ADMM;
        // columns that share a solver group (ADMM kernels) or a warp (active-set kernel: lock-step sweeps)
        const int width = plan->desc.precision == LIPM_PRECISION_F16X3 ? kF16Cols : (active_set ? 16 : 8);
        StreamBuffer order, unit_order;        // released on every path out of this scope
        if (B > 8 && (io->n_steps || (io->impulse && io->force_tick))) {
            rc = build_processing_order(B, n_max, io->n_steps, io->impulse, io->force_tick, width, order, stream);
            if (rc != LIPM_OK) return rc;
        }
        cb.order = order.as<int>();
        // longest-expected-first dispatch of the kernel's work units (8 columns FP64, 128 columns reduced precision)
        const int units = (cb.ncols + width - 1) / width;
        if (units > 1 && io->impulse && io->force_tick && !active_set) {
            rc = build_unit_order(units, width, cb.ncols, B, cb.order, n_max, io->n_steps, io->impulse, io->force_tick, unit_order, stream);
            if (rc != LIPM_OK) return rc;
        }
        cb.unit_order = unit_order.as<int>();
        rc = launch_strict(plan, cb, stream);
    }
    if (rc != LIPM_OK) return rc;
    if (io->outcome) {
        const double fall = io->fall_lean > 0 ? io->fall_lean : plan->desc.h * 0.57735026918962576451;   // h tan(30 deg)
        rc = launch_pack_outcomes(B, cb.max_violation, cb.lean_max, cb.status, cb.iters, fall, io->outcome, stream);
    }
    return rc;
}

int lipm_solve_tick(lipm_plan_t* plan, int mode, int B, const double* state, const double* z_min,
                    const double* z_max, int bounds_batch, double* next_state, double* zmp_plan,
                    double* jerk_plan, int32_t* status, int32_t* iters, void* stream) {
    if (!plan) { set_error("lipm_solve_tick: null plan"); return LIPM_E_INVALID; }
    if (mode != LIPM_MODE_STRICT && mode != LIPM_MODE_UNCONSTRAINED) { set_error("lipm_solve_tick: bad mode %d", mode); return LIPM_E_INVALID; }
    if (B < 0) { set_error("lipm_solve_tick: negative batch"); return LIPM_E_INVALID; }
    if (B == 0) return LIPM_OK;
    if (!state || !z_min || !z_max || !next_state) { set_error("lipm_solve_tick: state, z_min, z_max, next_state are required"); return LIPM_E_INVALID; }
    if (bounds_batch != 1 && bounds_batch != B) { set_error("lipm_solve_tick: bounds_batch must be 1 or B"); return LIPM_E_INVALID; }
    if (jerk_plan && !zmp_plan) { set_error("lipm_solve_tick: jerk_plan needs zmp_plan as well"); return LIPM_E_INVALID; }
    DeviceGuard guard(plan->desc.device);
    const int N = plan->N;

    // A single tick is a "walk" whose bound series is the window itself (rows 0..N-1, no bias)
    // and whose only produced sample (index 1) lands in next_state.  The per-axis [B,2] outputs
    // of the walk kernels are not used here: columns are the B problems themselves.
    ColumnBatch cb;
    std::memset(&cb, 0, sizeof(cb));
    cb.ncols = B;
    cb.B = B;                    // every column maps to "axis 0"
    cb.bounds_batch = bounds_batch;
    cb.bound_scn_stride = N;
    cb.bound_stride = 1;
    cb.bound_axis_offset = 0;
    cb.bound_index_bias = 0;
    cb.n_max = N;
    cb.n_steps = nullptr;
    cb.n_uniform = N;
    cb.ticks_override = 1;
    cb.init[0] = state; cb.init[1] = state;
    cb.z_min = z_min; cb.z_max = z_max;
    cb.hist[0] = next_state; cb.hist[1] = nullptr;
    cb.write_initial = 0;
    cb.hist_scn_stride = 3;
    cb.hist_index_bias = 1;
    cb.zmp_plan = (mode == LIPM_MODE_STRICT) ? zmp_plan : nullptr;

    // status / iters of the walk kernels are laid out [B,2]; stage them when requested
    StreamBuffer stage_buf;
    int32_t* stage = nullptr;
    if (status || iters) {
        LIPM_CUDA_TRY(stage_buf.alloc(sizeof(int32_t) * 4 * (size_t)B, (cudaStream_t)stream));
        stage = stage_buf.as<int32_t>();
        cb.status = stage;
        cb.iters = stage + 2 * (size_t)B;
    }
    int rc = mode != LIPM_MODE_STRICT ? launch_unconstrained_rollout(plan, cb, (cudaStream_t)stream)
             : launch_strict(plan, cb, (cudaStream_t)stream);
    if (rc == LIPM_OK && stage) {
        if (status) LIPM_CUDA_TRY(cudaMemcpy2DAsync(status, sizeof(int32_t), stage, 2 * sizeof(int32_t), sizeof(int32_t), B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        if (iters) LIPM_CUDA_TRY(cudaMemcpy2DAsync(iters, sizeof(int32_t), stage + 2 * (size_t)B, 2 * sizeof(int32_t), sizeof(int32_t), B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    if (rc != LIPM_OK) return rc;

    if (zmp_plan && mode == LIPM_MODE_UNCONSTRAINED) {
        rc = launch_unconstrained_plan(plan, B, state, z_min, z_max, bounds_batch, zmp_plan, (cudaStream_t)stream);
        if (rc != LIPM_OK) return rc;
    }
    if (jerk_plan) rc = launch_jerk_from_zmp(plan, B, state, zmp_plan, jerk_plan, (cudaStream_t)stream);
    return rc;
}

}  // extern "C"
