// Internal declarations shared by the translation units of liblipm_mpc.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "lipm_mpc.h"

namespace lipm {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
#define LIPM_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            ::lipm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                              __FILE__, __LINE__);                                       \
            return LIPM_E_CUDA;                                                          \
        }                                                                                \
    } while (0)

// Stream-ordered scratch memory that is released on every exit path (cudaFreeAsync on the stream it
// was allocated on), so an early LIPM_CUDA_TRY return does not leak it.
struct StreamBuffer {
    void* ptr = nullptr;
    cudaStream_t stream = nullptr;
    StreamBuffer() = default;
    StreamBuffer(const StreamBuffer&) = delete;
    StreamBuffer& operator=(const StreamBuffer&) = delete;
    ~StreamBuffer() { reset(); }
    void reset() {
        if (ptr) cudaFreeAsync(ptr, stream);
        ptr = nullptr;
    }
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        reset();
        stream = s;
        return cudaMallocAsync(&ptr, bytes ? bytes : 16, s);
    }
    cudaError_t alloc_zero(size_t bytes, cudaStream_t s) {
        cudaError_t e = alloc(bytes, s);
        return e != cudaSuccess ? e : cudaMemsetAsync(ptr, 0, bytes ? bytes : 16, s);
    }
    template <class T> T* as() const { return static_cast<T*>(ptr); }
    void* release() { void* p = ptr; ptr = nullptr; return p; }
};

// Per-launch scheduler state of the persistent kernels: [0..1] work counter (+pad), [2..9] four 64-bit statistics
// words.  It lives in stream-ordered memory of ITS launch, so rollouts on one plan may overlap on
// different streams; the statistics of the launch are copied into the plan afterwards (lipm_plan_stats).
constexpr size_t kLaunchStateBytes = 16 + 4 * sizeof(unsigned long long);
struct LaunchState {
    StreamBuffer buf;
    cudaError_t begin(cudaStream_t s) { return buf.alloc_zero(kLaunchStateBytes, s); }
    unsigned int* counter() const { return buf.as<unsigned int>(); }
    unsigned long long* stats() const { return reinterpret_cast<unsigned long long*>(buf.as<unsigned char>() + 16); }
    // file the statistics of this launch in the plan (last writer wins when launches overlap)
    cudaError_t publish(unsigned long long* plan_stats) const {
        return cudaMemcpyAsync(plan_stats, stats(), 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, buf.stream);
    }
};

// ------------------------------------------------------------------ geometry helpers
// DMMA.8x8x4 tiles: rows and the contraction index are padded to a multiple of 8.
__host__ __device__ inline int padded8(int n) { return (n + 7) & ~7; }
// Row pitch (in doubles) of the per-column vectors kept in shared memory as the B operand
// [column][k].  pitch % 16 is 4 or 12, which makes the B-fragment load (lane -> column
// lane/4, k lane%4) hit 16 distinct 8-byte banks per half warp.
__host__ __device__ inline int column_pitch(int np) { return np + 4; }

constexpr int kRowGroups = 4;       // warps along the row dimension of the iteration matrix
constexpr int kMaxRowBlocksPerWarp = 5;
constexpr int kMaxResidentHorizon = 8 * kRowGroups * kMaxRowBlocksPerWarp;  // 160: matrix resident in shared memory
constexpr int kMaxStrictHorizon = 8 * 16 * kMaxRowBlocksPerWarp;             // 640: matrix streamed from L2 beyond that

// Reduced-precision strict kernel (tcgen05): fixed 128 x 160 x 160 tile.  Rows 0..151 carry the
// horizon, contraction slots 152..154 carry the axis state (the G Px x term rides in the K padding).
constexpr int kF16Tile = 160;                 // padded horizon = MMA N = MMA K
constexpr int kF16Cols = 128;                 // columns (scenario axes) per CTA = MMA M = TMEM lanes
constexpr int kF16StateSlot = 152;            // first contraction slot of the state
constexpr int kMaxReducedHorizon = 152;         // resident operands (lipm_rollout_strict_f16.cu)
constexpr int kMaxReducedStreamK = 640;         // streamed operands (lipm_rollout_strict_f16s.cu): padded contraction length
constexpr int kMaxReducedStreamHorizon = kMaxReducedStreamK - 3;   // 637: three contraction slots carry the axis state
constexpr int kF16OperandBytesB = (kF16Tile / 8) * (kF16Tile / 8) * 128;   // 51200: one FP16 image of B
constexpr int kF16OperandBytesA = (kF16Tile / 8) * (kF16Cols / 8) * 128;   // 40960: one FP16 image of A
// Byte offset of element (row, k) of a K-major operand without swizzle: 8 x 16-byte core matrices,
// 128 B between 8-row groups (SBO), rows/8 * 128 B between 8-element K chunks (LBO).
__host__ __device__ inline int f16_operand_offset(int row, int k, int rows) {
    return (k >> 3) * ((rows >> 3) * 128) + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

// ------------------------------------------------------------------ the plan
}  // namespace lipm

struct lipm_plan {
    lipm_plan_desc_t desc;
    int N;        // horizon
    int NP;       // padded8(N)
    int sm_count;
    // model scalars (host copies)
    double T, hg;         // dt, h/g
    double pu00;          // Pu[0,0] = T^3/6 - T h/g
    // device matrices (all float64)
    double* d_px;         // [N,3]
    double* d_pu;         // [N,N] dense lower-triangular Toeplitz
    double* d_abc;        // [15]
    double* d_tinv;       // [N]   first column of Pu^-1 (Pu^-1 is lower-triangular Toeplitz)
    double* d_g;          // [N,N] G = R Pu^-T Pu^-1
    double* d_gpx;        // [NP,3] G Px, zero rows beyond N
    double* d_work;       // [N,N] factorisation scratch
    double* d_hinv;       // [N,N] (Q I + G)^-1
    double* d_minv;       // [N,N] (Q I + G + rho I)^-1
    double* d_mfrag;      // [(NP/8) x (NP/4) x 32] ADMM matrix in DMMA A-fragment order
    double* d_hfrag;      // same layout: H^-1 (dense unconstrained preview on the FP64 tensor pipe)
    double* d_gain;       // [N+3] kz, kx
    unsigned char* d_bf16; // [2 x kF16OperandBytesB] FP16 hi / lo images of [M | M G Px] (F16X3 plans, N <= 152)
    unsigned char* d_f16s; // FP16 hi / lo images of [M | M G Px] cut into 160-row slabs (F16X3 plans, N > 152)
    int f16s_nk;           // padded contraction length of d_f16s: multiple of 16, >= N + 3
    int f16s_slabs;        // 160-row slabs
    int single_cta_f16;   // 1: reduced-precision walks on the single-CTA kernel instead of CTA pairs (lipm_set_option)
    int solver;           // LIPM_SOLVER_* of the FP64 strict mode (lipm_set_option; 0 = active set + Riccati sweeps)
    int riccati_lanes;    // columns per warp of the active-set kernel (0 = automatic)
    int riccati_max_sweeps;   // active-set corrections per QP (0 -> 64)
    int riccati_few_pins; // pinned rows up to which a sweep uses table solve + dense correction (0 = never, the default: measured slower in a one-thread-per-column kernel)
    int riccati_scratch;  // feed-forward terms of the table stages: 0 automatic, 1 shared memory, 2 global scratch
    double* d_ric_tab;    // [N + 4][16] per-stage tables of the active-set kernel's free tail (build_riccati_tables)
    double* d_ric_aux;    // [N + 1][12] K_j, P_j of the same stages
    double* d_kx;         // [N][3] Kx = H^-1 G Px: state part of the unconstrained plan (row-parallel kernel, N <= 160)
    int rowpar_chunks;    // slot chunks of an active-set rollout on side streams (LIPM_OPT_WALK_CHUNKS): 0 automatic, 1 none, 2..4
    int rowpar_off;       // 1: FP64 active-set walks skip the row-parallel stage; -1: they use it whatever the batch size; 0: automatic
    int force_segments;   // >0: tick ranges per unit of the FP64 strict kernel (lipm_set_option; 0 = automatic)
    int* d_info;          // [4] factorisation status words
    unsigned long long* d_stats;   // [0] group iterations, [1] units of the last strict launch (copied from the launch's own state)
};

namespace lipm {

// plan construction (lipm_plan.cu)
int build_plan_matrices(lipm_plan* plan);
int build_riccati_tables(lipm_plan* plan);        // lipm_rollout_riccati.cu
int build_rowpar_tables(lipm_plan* plan);         // lipm_rollout_rowpar.cu

// Row-parallel first stage of the active-set solver (lipm_rollout_rowpar.cu) and its hand-over to k_riccati_rollout.
constexpr int kRowparWords = 5;        // 32-row words of a column: horizons up to 160
#ifndef LIPM_ROWPAR_MAX_PINS
#define LIPM_ROWPAR_MAX_PINS 16       // (at most 32: one lane per row of the small system)
#endif
constexpr int kRowparMaxPins = LIPM_ROWPAR_MAX_PINS;     // pinned rows a row-parallel sweep can take
struct ParkRecord {                    // a column that left the row-parallel kernel: where and in which state
    int tick, sweeps, notconv;
    int stage;                         // which launch of the sequence resumes it next (0: none; kept by column, so that the
                                       // Riccati kernel finds its columns in processing order: similar walks share a warp)
    double x[3];
    double viol, lean;
    uint32_t m_lo[kRowparWords], m_hi[kRowparWords];
};
constexpr int kRowparLists = 3;        // hand-over lists of a chunk: row-parallel -> Riccati -> row-parallel -> Riccati
constexpr int kRowparMaxChunks = 4;    // slot chunks of a rollout, each with its own hand-over lists (and side stream)
constexpr int kRowparAutoChunks = 4;   // ... chosen automatically for batches of 4096 scenarios or more
struct RowparScratch {
    StreamBuffer buf;
    double* u = nullptr;               // shared preview terms
    double2* pk = nullptr;             // bound pairs per (series, axis)
    int* n_series = nullptr;
    ParkRecord* park = nullptr;        // [ncols]
    int* lists = nullptr;              // [chunks][kRowparLists][ncols] parked columns
    int* counts = nullptr;             // [chunks][4]
    int ncols = 0;
    unsigned long long* stats = nullptr;
    int* list(int chunk, int l) const { return lists + ((size_t)chunk * kRowparLists + l) * ncols; }
    int* count(int chunk, int l) const { return counts + chunk * 4 + l; }
    static int stage(int chunk, int l) { return 1 + chunk * 4 + l; }   // stage number in the records of list l of a chunk
};

// One "column" = one axis of one scenario.  The rollout kernels receive this flat description.
struct ColumnBatch {
    int ncols;              // number of columns (2B for walks, B for single ticks)
    int B;                  // scenarios; column c -> slot c % B, axis c / B
    const int* order;       // [B] scenario processed in each slot (nullptr: identity)
    const int* unit_order;  // [units] dispatch order of the kernel's work units (nullptr: last unit first)
    int bounds_batch;       // 1 (shared), B (one series per scenario) or V (with bounds_index)
    const int32_t* bounds_index;  // [B] series used by each scenario, or nullptr
    long long bound_scn_stride;   // doubles between consecutive scenarios' bound series
    int bound_stride;       // doubles between consecutive samples of one axis (2 walks, 1 ticks)
    int bound_axis_offset;  // doubles between the x and the y series (1 for walks)
    int bound_index_bias;   // window row j of tick i reads sample min(i + 1 + j, n-1) - bias
    int n_max;              // allocated samples per scenario in the histories
    const int32_t* n_steps; // [B] or nullptr (-> n_uniform)
    int n_uniform;
    int ticks_override;     // >0: run exactly this many ticks (single-tick API); else n-1
    const double* init[2];  // x0, y0 ([B,3])
    const double* z_min;
    const double* z_max;
    const double* impulse;  // [B] or nullptr
    const int32_t* force_tick;  // [B] or nullptr
    double* hist[2];        // x_hist, y_hist or nullptr
    int write_initial;      // store sample 0 (walks) or not (single ticks)
    long long hist_scn_stride;  // doubles between scenarios in hist (n_max*3 walks, 3 ticks)
    int hist_index_bias;    // sample s is stored at row s - bias
    uint8_t* flags;         // [B,n_max,2] or nullptr
    double* max_violation;  // [B,2] or nullptr
    int32_t* status;        // [B,2] or nullptr
    int32_t* iters;         // [B,2] or nullptr
    double* zmp_plan;       // [ncols,N] or nullptr: final previewed ZMP of the last tick
    // read-out fused into the walk kernels (the drivers' C . state, scripts/run_mpc.py:294)
    double* com;            // [B,n_max,2] CoM position per sample (x, y interleaved) or nullptr
    double* zmp;            // [B,n_max,2] realised ZMP  C . state  per sample or nullptr
    double* lean_max;       // [B,2] per axis: max over samples of |CoM - ZMP| or nullptr
};

int build_processing_order(int B, int n_uniform, const int32_t* n_steps, const double* impulse,
                           const int32_t* force_tick, int width, StreamBuffer& out, cudaStream_t stream);
int build_unit_order(int num_units, int width, int ncols, int B, const int* order, int n_uniform, const int32_t* n_steps,
                     const double* impulse, const int32_t* force_tick, StreamBuffer& out, cudaStream_t stream);
// per-scenario outcome record from the per-axis results of a walk (lipm_rollout_ex)
int launch_pack_outcomes(int B, const double* max_violation, const double* lean_max, const int32_t* status, const int32_t* iters,
                         double fall_lean, lipm_outcome_t* outcome, cudaStream_t stream);
int launch_strict_rollout(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);
int launch_strict_rollout_riccati(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);   // active set: row-parallel stage, then Riccati sweeps
bool rowpar_eligible(const lipm_plan* plan, const ColumnBatch& cb);
int rowpar_prepare(lipm_plan* plan, const ColumnBatch& cb, RowparScratch& rs, cudaStream_t stream);
int launch_rowpar(lipm_plan* plan, const ColumnBatch& cb, const RowparScratch& rs, int chunk, int chunks, int in_list, int out_list,
                  cudaStream_t stream);
int launch_strict_rollout_f16(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);
int launch_strict_rollout_f16s(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);
int launch_strict_rollout_f16p(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);   // CTA pairs (cta_group::2)
// strict rollout on the kernel that matches the plan's precision and horizon
inline int launch_strict(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    if (plan->desc.precision != LIPM_PRECISION_F16X3)
        return plan->solver == LIPM_SOLVER_ADMM ? launch_strict_rollout(plan, cb, stream) : launch_strict_rollout_riccati(plan, cb, stream);
    if (plan->N > kMaxReducedHorizon) return launch_strict_rollout_f16s(plan, cb, stream);
    return plan->single_cta_f16 ? launch_strict_rollout_f16(plan, cb, stream) : launch_strict_rollout_f16p(plan, cb, stream);
}
int launch_unconstrained_rollout(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream);
int launch_jerk_from_zmp(lipm_plan* plan, int ncols, const double* state, const double* zmp_plan,
                         double* jerk_plan, cudaStream_t stream);
int launch_unconstrained_plan(lipm_plan* plan, int ncols, const double* state, const double* z_min,
                              const double* z_max, int bounds_batch, double* zmp_plan,
                              cudaStream_t stream);

}  // namespace lipm
