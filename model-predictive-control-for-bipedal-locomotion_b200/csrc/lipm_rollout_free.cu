// Unconstrained ("non-strict") mode and the small dense helpers of the single-tick API.
//
// Reference: predict_wieber_axis, non-strict branch (zmp_controller.py:196-199)
//     X = -inv(Pu'Pu + R/Q I) Pu' (Px x - zref);   next = A x + B X[0]
// rebuilt (O(N^2) Python loop) and re-inverted (O(N^3)) on every call.  Only X[0] is consumed, and
//     X[0] = kz . zref_window + kx . x
// with a gain row (kz, kx) that depends on the plan only.  A walk is therefore a sliding
// correlation of the zref series with kz -- independent of the state, so computed once per bound
// series (k_free_preview), not once per scenario -- followed by the 3-state recursion, one thread
// per column (k_free_rollout).  Bandwidth bound: per solve one (z_min, z_max) sample is read (16 B,
// shared through L1 when the series is) and the state (24 B) and a flag are written.
#include "lipm_internal.cuh"
#include "lipm_device.cuh"

namespace lipm {

namespace {

constexpr int kFreeWarps = 4;     // warps per CTA
constexpr int kFreeCols = 8;      // columns per warp (lanes 0..7 run the recursions, all 32 lanes write out): the recursion is a
                                  // latency chain, so a sweep needs many warps in flight more than it needs full warps
constexpr int kFreeTicks = 8;     // ticks staged in shared memory before a column's states are written out
constexpr int kFreeChunk = 65535; // columns per launch pair when every column needs its own preview series

struct FreeParams {
    ColumnBatch cb;
    int N;
    double viol_tol;
    double T, c2, c3, hg;
    const double* gain;   // kz[N], kx[3]
    double* preview;      // [series][ticks_max] state-independent part of the first jerk, kz . zref_window
    int ticks_max;
    int per_column;       // 1: one preview series per column; 0: one per (bound series, axis)
    int col0, ncols;      // columns of this launch
    int pairs;            // 1 (walks): column c = axis c & 1 of scenario c >> 1, so that the x and y lanes of a scenario are
                          // neighbours and its com / zmp / flag samples leave as contiguous runs; 0 (single ticks): c -> slot c
};

struct ColumnView {
    int axis, b, n, ticks, last;
    long long bz;              // bound series of the scenario
    const double *plo, *phi;
};

// Decodes column c: like the strict kernels do (axis-major) or with the axes of a scenario interleaved.
__device__ __forceinline__ ColumnView column_view(const ColumnBatch& B, int c, int pairs) {
    ColumnView v;
    v.axis = pairs ? (c & 1) : c / B.B;
    v.b = pairs ? (c >> 1) : c - v.axis * B.B;
    int n = B.n_steps ? B.n_steps[v.b] : B.n_uniform;
    n = n > B.n_max ? B.n_max : n;
    int ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
    if (n < 1 || ticks < 0) ticks = 0;
    v.n = n; v.ticks = ticks; v.last = n - 1;
    v.bz = B.bounds_index ? B.bounds_index[v.b] : ((B.bounds_batch == 1) ? 0 : v.b);
    v.plo = B.z_min + v.bz * B.bound_scn_stride + v.axis * B.bound_axis_offset;
    v.phi = B.z_max + v.bz * B.bound_scn_stride + v.axis * B.bound_axis_offset;
    return v;
}

// preview[series][i] = sum_j kz[j] * zref[min(i + bias + j, last)]: the part of the first jerk that does not depend
// on the state.  One thread per (series, tick); scenarios that share their bound series share the result, so a
// sweep over pushes or initial states computes it twice (x and y), not 2 B times.
__global__ void __launch_bounds__(128) k_free_preview(const FreeParams P) {
    extern __shared__ double sK[];
    const ColumnBatch& B = P.cb;
    const int N = P.N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) sK[j] = P.gain[j];
    __syncthreads();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nseries = P.per_column ? P.ncols : B.bounds_batch * 2;
    if (gid >= (long long)nseries * P.ticks_max) return;
    const int series = (int)(gid / P.ticks_max), i = (int)(gid - (long long)series * P.ticks_max);
    const double *plo, *phi;
    int last, ticks;
    if (P.per_column) {
        const ColumnView v = column_view(B, P.col0 + series, P.pairs);
        plo = v.plo; phi = v.phi; last = v.last; ticks = v.ticks;
    } else {
        const int axis = series & 1;
        const long long bz = series >> 1;
        plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        int n = B.n_uniform > B.n_max ? B.n_max : B.n_uniform;
        ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
        if (n < 1 || ticks < 0) ticks = 0;
        last = n - 1;
    }
    if (i >= ticks) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int s0 = i + B.bound_index_bias;
    int j = 0;
    for (; j + 4 <= N; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int s = s0 + j + u;
            s = s > last ? last : s;
            const double z = (__ldg(plo + (long long)s * B.bound_stride) + __ldg(phi + (long long)s * B.bound_stride)) / 2;
            acc[u] = fma(sK[j + u], z, acc[u]);
        }
    }
    for (; j < N; ++j) {
        int s = s0 + j;
        s = s > last ? last : s;
        const double z = (__ldg(plo + (long long)s * B.bound_stride) + __ldg(phi + (long long)s * B.bound_stride)) / 2;
        acc[j & 3] = fma(sK[j], z, acc[j & 3]);
    }
    P.preview[(size_t)series * P.ticks_max + i] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// The same for walks (many ticks per series): blockIdx.y = series, a block covers 128 consecutive ticks and stages
// the N + 127 zref samples they share in shared memory.  Same summation order as k_free_preview, same bits.
__global__ void __launch_bounds__(128) k_free_preview_walk(const FreeParams P) {
    extern __shared__ double sm[];
    const ColumnBatch& B = P.cb;
    const int N = P.N;
    double* sK = sm;
    double* sZ = sm + N;
    const int series = blockIdx.y, i0 = blockIdx.x * 128;
    const double *plo, *phi;
    int last, ticks;
    if (P.per_column) {
        const ColumnView v = column_view(B, P.col0 + series, P.pairs);
        plo = v.plo; phi = v.phi; last = v.last; ticks = v.ticks;
    } else {
        const int axis = series & 1;
        const long long bz = series >> 1;
        plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
        int n = B.n_uniform > B.n_max ? B.n_max : B.n_uniform;
        ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
        if (n < 1 || ticks < 0) ticks = 0;
        last = n - 1;
    }
    if (i0 >= ticks) return;
    for (int j = threadIdx.x; j < N; j += 128) sK[j] = P.gain[j];
    for (int k = threadIdx.x; k < N + 127; k += 128) {
        int s = i0 + B.bound_index_bias + k;
        s = s > last ? last : s;
        sZ[k] = (__ldg(plo + (long long)s * B.bound_stride) + __ldg(phi + (long long)s * B.bound_stride)) / 2;
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= ticks) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double* z = sZ + threadIdx.x;
    int j = 0;
    for (; j + 4 <= N; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(sK[j + u], z[j + u], acc[u]);
    }
    for (; j < N; ++j) acc[j & 3] = fma(sK[j], z[j], acc[j & 3]);
    P.preview[(size_t)series * P.ticks_max + i] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// The 3-state recursion of a walk, one THREAD per column (kFreeCols columns per warp): x <- A x + B (preview[i] + kx . x), push, box test.
// The states of kFreeTicks ticks are staged in shared memory and written out per column as one contiguous run.
__global__ void __launch_bounds__(32 * kFreeWarps) k_free_rollout(const FreeParams P) {
    __shared__ double tile[kFreeWarps][kFreeCols][kFreeTicks * 3 + 1];
    __shared__ unsigned long long flags8[kFreeWarps][kFreeCols];     // one byte per staged tick
    __shared__ double tile2[kFreeWarps][2][kFreeCols / 2][kFreeTicks * 2];   // com, zmp of 4 scenarios: (x, y) per staged tick
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ColumnBatch& B = P.cb;
    const int N = P.N;
    const double kx0 = P.gain[N], kx1 = P.gain[N + 1], kx2 = P.gain[N + 2];

    const int local = (blockIdx.x * kFreeWarps + warp) * kFreeCols + lane;      // column within this launch
    const bool valid = lane < kFreeCols && local < P.ncols;
    const int c = P.col0 + (valid ? local : 0);
    const ColumnView v = column_view(B, c, P.pairs);
    const int ticks = valid ? v.ticks : 0;
    double* hist = nullptr;
    double *pcom = nullptr, *pzmp = nullptr;
    uint8_t* pf = nullptr;
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, imp = 0.0, viol = 0.0, lean = 0.0;
    int ftick = -1;
    const double* preview = P.preview + (size_t)(P.per_column ? (long long)local : v.bz * 2 + v.axis) * P.ticks_max;
    if (valid) {
        hist = v.axis ? B.hist[1] : B.hist[0];
        if (hist) hist += (long long)v.b * B.hist_scn_stride;
        pf = B.flags ? B.flags + ((long long)v.b * B.n_max) * 2 + v.axis : nullptr;
        pcom = B.com ? B.com + ((long long)v.b * B.n_max) * 2 + v.axis : nullptr;
        pzmp = B.zmp ? B.zmp + ((long long)v.b * B.n_max) * 2 + v.axis : nullptr;
        const double* init = (v.axis ? B.init[1] : B.init[0]) + (long long)v.b * 3;
        x0 = init[0]; x1 = init[1]; x2 = init[2];
        if (v.axis == 1 && B.impulse != nullptr) {
            imp = B.impulse[v.b];
            ftick = B.force_tick ? B.force_tick[v.b] : -1;
        }
        if (B.write_initial && ticks > 0) {
            const double zmp = x0 - P.hg * x2;
            const double below = v.plo[0] - zmp, above = zmp - v.phi[0];
            viol = fmax(fmax(below, above), 0.0);
            lean = fabs(x0 - zmp);
            if (hist) { hist[0] = x0; hist[1] = x1; hist[2] = x2; }
            if (pcom) pcom[0] = x0;
            if (pzmp) pzmp[0] = zmp;
            if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                      (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
        }
    }
    int warp_ticks = ticks;
    for (int o = 16; o; o >>= 1) warp_ticks = max(warp_ticks, __shfl_xor_sync(0xffffffffu, warp_ticks, o));

    for (int i0 = 0; i0 < warp_ticks; i0 += kFreeTicks) {
        // everything the block's steps read from memory first (independent of the recursion), then the
        // recursion in registers
        double pv[kFreeTicks], l0[kFreeTicks], h0[kFreeTicks];
#pragma unroll
        for (int l = 0; l < kFreeTicks; ++l) {
            const int i = i0 + l;
            const bool on = i < ticks;
            int s_row0 = i + B.bound_index_bias;
            s_row0 = s_row0 > v.last ? v.last : s_row0;
            s_row0 = s_row0 < 0 ? 0 : s_row0;
            pv[l] = on ? __ldg(preview + i) : 0.0;
            l0[l] = on ? __ldg(v.plo + (long long)s_row0 * B.bound_stride) : 0.0;
            h0[l] = on ? __ldg(v.phi + (long long)s_row0 * B.bound_stride) : 0.0;
        }
        unsigned long long flag8 = 0ull;
#pragma unroll
        for (int l = 0; l < kFreeTicks; ++l) {
            const int i = i0 + l;
            if (i < ticks) {
                const double jerk = pv[l] + fma(kx2, x2, fma(kx1, x1, kx0 * x0));
                const double n0 = fma(P.c3, jerk, fma(P.c2, x2, fma(P.T, x1, x0)));
                double n1 = fma(P.c2, jerk, fma(P.T, x2, x1));
                const double n2 = fma(P.T, jerk, x2);
                if (i == ftick) n1 -= imp;
                x0 = n0; x1 = n1; x2 = n2;
                tile[warp][lane][l * 3 + 0] = n0; tile[warp][lane][l * 3 + 1] = n1; tile[warp][lane][l * 3 + 2] = n2;   // lane < kFreeCols here
                const double zmp = n0 - P.hg * n2;
                const double below = l0[l] - zmp, above = zmp - h0[l];
                viol = fmax(viol, fmax(fmax(below, above), 0.0));
                lean = fmax(lean, fabs(n0 - zmp));
                if (P.pairs) {       // staged: the (x, y) lanes of a scenario fill 16-byte pairs, written out as one run below
                    tile2[warp][0][lane >> 1][l * 2 + (lane & 1)] = n0;
                    tile2[warp][1][lane >> 1][l * 2 + (lane & 1)] = zmp;
                } else {
                    if (pcom) pcom[(long long)(i + 1) * 2] = n0;
                    if (pzmp) pzmp[(long long)(i + 1) * 2] = zmp;
                }
                flag8 |= (unsigned long long)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                              (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u)) << (8 * l);
            }
        }
        if (lane < kFreeCols) flags8[warp][lane] = flag8;
        __syncwarp();
        // write-out: column by column the run of this block's states (up to 8 x 24 bytes, contiguous), four columns
        // at a time their flags (8 bytes each at stride 2: x and y flags of a sample are neighbours)
        int run = ticks - i0;
        run = run < 0 ? 0 : (run > kFreeTicks ? kFreeTicks : run);
        const unsigned long long base = hist ? (unsigned long long)(hist + ((long long)(i0 + 1) - B.hist_index_bias) * 3) : 0ull;
        const unsigned long long fbase = pf ? (unsigned long long)(pf + (long long)(i0 + 1) * 2) : 0ull;
#pragma unroll
        for (int col = 0; col < kFreeCols; ++col) {
            const int nv = __shfl_sync(0xffffffffu, run, col) * 3;
            const unsigned long long dst = __shfl_sync(0xffffffffu, base, col);
            if (dst != 0ull && lane < nv) reinterpret_cast<double*>(dst)[lane] = tile[warp][col][lane];
        }
        if (P.pairs && (B.com || B.zmp)) {
            // com / zmp: per scenario the (x, y) samples of this block's ticks are 16 * run contiguous bytes
            const unsigned long long cbase = pcom ? (unsigned long long)(pcom + (long long)(i0 + 1) * 2) : 0ull;
            const unsigned long long zbase = pzmp ? (unsigned long long)(pzmp + (long long)(i0 + 1) * 2) : 0ull;
#pragma unroll
            for (int sc = 0; sc < kFreeCols / 2; ++sc) {
                const int nv = __shfl_sync(0xffffffffu, run, 2 * sc) * 2;          // x lane of the scenario (axis 0: pointers have no axis offset)
                const unsigned long long dc = __shfl_sync(0xffffffffu, cbase, 2 * sc);
                const unsigned long long dz = __shfl_sync(0xffffffffu, zbase, 2 * sc);
                const int k = lane & 15;
                const unsigned long long dst = lane < 16 ? dc : dz;
                if (dst != 0ull && k < nv) reinterpret_cast<double*>(dst)[k] = tile2[warp][lane >> 4][sc][k];
            }
        }
#pragma unroll
        for (int q = 0; q < kFreeCols / 4; ++q) {
            const int col = q * 4 + (lane >> 3), l = lane & 7;
            const int nv = __shfl_sync(0xffffffffu, run, col);
            const unsigned long long dst = __shfl_sync(0xffffffffu, fbase, col);
            if (dst != 0ull && l < nv) reinterpret_cast<uint8_t*>(dst)[l * 2] = (uint8_t)(flags8[warp][col] >> (8 * l));
        }
        __syncwarp();
    }
    if (valid) {
        if (B.max_violation) B.max_violation[v.b * 2 + v.axis] = viol;
        if (B.lean_max) B.lean_max[v.b * 2 + v.axis] = lean;
        if (B.status) B.status[v.b * 2 + v.axis] = 0;
        if (B.iters) B.iters[v.b * 2 + v.axis] = 0;
    }
}

// Full unconstrained preview of a batch: Z = H^-1 (Q Zref + G Px X), one dense [N x N] x [N x columns] product
// on the FP64 tensor pipe (north_star: "a fixed factor followed by a batched solve over B right-hand sides";
// H^-1 is formed once per plan from the Cholesky factor, so the batched solve is this product; 2 N^2 flops
// per column).  Same organisation as the strict kernel: 4-warp groups of 8 columns, H^-1 in DMMA
// fragment order -- resident in shared memory up to N = 160, read from L2 beyond.
struct DenseParams {
    int N, NP, ncols, bounds_batch;
    double Q;
    const double *hfrag, *gpx, *state, *z_min, *z_max;
    double* zmp_plan;
    unsigned int* tile_counter;
    unsigned int num_units;
};

template <int NRB, int NG, bool STREAM>
__global__ void __launch_bounds__(32 * kRowGroups * NG, 1) k_free_dense(const DenseParams P) {
    constexpr int kGroupThreads = 32 * kRowGroups;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = P.N, NP = P.NP;
    const int KB = NP / 4, RBLK = NP / 8, KP = NP + 4;
    const int aStride = KB * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / kRowGroups, gtid = tid & (kGroupThreads - 1);
    const int rg = warp & (kRowGroups - 1);
    const int g = lane >> 2, t = lane & 3;
    const int bar_id = 1 + grp;
    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(smem_raw);
    int* sUnit = reinterpret_cast<int*>(smem_raw + 16) + grp;
    double* sR = reinterpret_cast<double*>(smem_raw + 64) + grp * (kGroupCols * KP);
    const double* sM = STREAM ? P.hfrag : reinterpret_cast<double*>(smem_raw + 64) + NG * (kGroupCols * KP);
    if (!STREAM) {
        if (tid == 0) mbar_init(sBar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t m_bytes = (uint32_t)((size_t)RBLK * KB * 32 * sizeof(double));
            mbar_expect_tx(sBar, m_bytes);
            unsigned char* dst = reinterpret_cast<unsigned char*>(reinterpret_cast<double*>(smem_raw + 64) + NG * (kGroupCols * KP));
            for (uint32_t off = 0; off < m_bytes; off += 32768) {
                const uint32_t nb = (m_bytes - off < 32768u) ? (m_bytes - off) : 32768u;
                bulk_load(dst + off, reinterpret_cast<const unsigned char*>(P.hfrag) + off, nb, sBar);
            }
        }
        mbar_wait(sBar, 0);
    }
    int nrb = RBLK - rg * NRB;
    nrb = nrb < 0 ? 0 : (nrb > NRB ? NRB : nrb);
    const int row0 = rg * NRB * 8 + g;
    const int lc0 = 2 * t;
    const double* aBase = sM + (size_t)(rg * NRB) * aStride + lane;
    const double* bBase = sR + g * KP + t;
    double* rOut0 = sR + lc0 * KP + row0;
    for (;;) {
        group_sync<kGroupThreads>(bar_id);
        if (gtid == 0) *sUnit = (int)atomicAdd(P.tile_counter, 1u);
        group_sync<kGroupThreads>(bar_id);
        const unsigned int unit = (unsigned int)*sUnit;
        if (unit >= P.num_units) break;
        const int col0 = (int)unit * kGroupCols;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = col0 + lc0 + e;
            const bool live = c < P.ncols;
            const long long bz = P.bounds_batch == 1 ? 0 : (live ? c : 0);
            const double x0 = live ? P.state[(size_t)c * 3] : 0.0, x1 = live ? P.state[(size_t)c * 3 + 1] : 0.0,
                         x2 = live ? P.state[(size_t)c * 3 + 2] : 0.0;
#pragma unroll
            for (int r = 0; r < NRB; ++r) {
                const int j = row0 + r * 8;
                double rhs = 0.0;
                if (live && r < nrb && j < N) {
                    const double zref = (P.z_max[bz * N + j] + P.z_min[bz * N + j]) / 2;
                    rhs = fma(P.Q, zref, fma(P.gpx[j * 3 + 2], x2, fma(P.gpx[j * 3 + 1], x1, P.gpx[j * 3] * x0)));
                }
                if (r < nrb) rOut0[e * KP + r * 8] = rhs;
            }
        }
        group_sync<kGroupThreads>(bar_id);
        double acc[NRB][2];
#pragma unroll
        for (int r = 0; r < NRB; ++r) acc[r][0] = acc[r][1] = 0.0;
        ContractDispatch<NRB, NRB, 0, STREAM>::run(nrb, acc, aBase, aStride, bBase, KB);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = col0 + lc0 + e;
            if (c >= P.ncols) continue;
#pragma unroll
            for (int r = 0; r < NRB; ++r) {
                const int j = row0 + r * 8;
                if (r < nrb && j < N) P.zmp_plan[(size_t)c * N + j] = acc[r][e];
            }
        }
    }
}

template <int NRB, int NG, bool STREAM>
int launch_dense(lipm_plan* plan, DenseParams& P, cudaStream_t stream) {
    const size_t nfrag = STREAM ? 0 : (size_t)(P.NP / 8) * (P.NP / 4) * 32;
    const size_t smem = 64 + sizeof(double) * ((size_t)NG * kGroupCols * column_pitch(P.NP) + nfrag);
    auto kernel = k_free_dense<NRB, NG, STREAM>;
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned ctas = (P.num_units + NG - 1) / NG;
    const unsigned grid = ctas < (unsigned)plan->sm_count ? ctas : (unsigned)plan->sm_count;
    LaunchState state;
    LIPM_CUDA_TRY(state.begin(stream));
    P.tile_counter = state.counter();
    kernel<<<grid, 32 * kRowGroups * NG, smem, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

// jerk_plan[c] = Pu^-1 (zmp_plan[c] - Px x): Pu^-1 is lower-triangular Toeplitz with first column t.
__global__ void k_jerk_from_zmp(int N, int ncols, const double* __restrict__ tinv,
                                const double* __restrict__ px, const double* __restrict__ state,
                                const double* __restrict__ zmp_plan, double* __restrict__ jerk_plan) {
    extern __shared__ double sh[];    // d[N] = zmp - Px x, then t[N]
    double* d = sh;
    double* tt = sh + N;
    const int c = blockIdx.x;
    if (c >= ncols) return;
    const double x0 = state[c * 3 + 0], x1 = state[c * 3 + 1], x2 = state[c * 3 + 2];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        d[j] = zmp_plan[(size_t)c * N + j] - fma(px[j * 3 + 2], x2, fma(px[j * 3 + 1], x1, px[j * 3 + 0] * x0));
        tt[j] = tinv[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc = fma(tt[i - j], d[j], acc);
        jerk_plan[(size_t)c * N + i] = acc;
    }
}

}  // namespace

int launch_unconstrained_rollout(lipm_plan* plan, const ColumnBatch& cb, cudaStream_t stream) {
    FreeParams P;
    P.cb = cb;
    P.N = plan->N;
    P.viol_tol = plan->desc.viol_tol;
    P.T = plan->T;
    P.c2 = std::pow(plan->desc.dt, 2) / 2.0;
    P.c3 = std::pow(plan->desc.dt, 3) / 6.0;
    P.hg = plan->hg;
    P.gain = plan->d_gain;
    if (cb.ncols <= 0) return LIPM_OK;
    P.ticks_max = cb.ticks_override > 0 ? cb.ticks_override : cb.n_max - 1;
    if (P.ticks_max <= 0) P.ticks_max = 1;
    // Scenarios that share a bound series (and its length) share the preview term; per-scenario lengths or series
    // need one preview series per column, computed and consumed in chunks so that the scratch stays small.
    P.per_column = (cb.n_steps != nullptr || (cb.bounds_batch != 1 && cb.bounds_index == nullptr)) ? 1 : 0;
    P.pairs = (cb.ncols == 2 * cb.B) ? 1 : 0;          // walks: both axes of every scenario (single ticks: ncols == B)
    const int chunk = P.per_column ? (kFreeChunk & ~1) : cb.ncols;      // even: a scenario's two columns stay in one launch
    const size_t series = P.per_column ? (size_t)(cb.ncols < chunk ? cb.ncols : chunk) : (size_t)cb.bounds_batch * 2;
    StreamBuffer preview;
    LIPM_CUDA_TRY(preview.alloc(sizeof(double) * series * (size_t)P.ticks_max, stream));
    P.preview = preview.as<double>();
    for (int col0 = 0; col0 < cb.ncols; col0 += chunk) {
        P.col0 = col0;
        P.ncols = cb.ncols - col0 < chunk ? cb.ncols - col0 : chunk;
        if (P.per_column || col0 == 0) {
            const int nseries = P.per_column ? P.ncols : cb.bounds_batch * 2;
            if (P.ticks_max >= 32 && nseries <= 65535) {
                k_free_preview_walk<<<dim3((unsigned)((P.ticks_max + 127) / 128), (unsigned)nseries), 128,
                                      sizeof(double) * (2 * (size_t)plan->N + 127), stream>>>(P);
            } else {
                const long long threads = (long long)nseries * P.ticks_max;
                k_free_preview<<<(unsigned)((threads + 127) / 128), 128, sizeof(double) * plan->N, stream>>>(P);
            }
        }
        const unsigned grid = (unsigned)((P.ncols + kFreeCols * kFreeWarps - 1) / (kFreeCols * kFreeWarps));
        k_free_rollout<<<grid, 32 * kFreeWarps, 0, stream>>>(P);
    }
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

int launch_unconstrained_plan(lipm_plan* plan, int ncols, const double* state, const double* z_min,
                              const double* z_max, int bounds_batch, double* zmp_plan, cudaStream_t stream) {
    if (ncols <= 0) return LIPM_OK;
    DenseParams P;
    P.N = plan->N; P.NP = plan->NP; P.ncols = ncols; P.bounds_batch = bounds_batch;
    P.Q = plan->desc.Q;
    P.hfrag = plan->d_hfrag; P.gpx = plan->d_gpx; P.state = state; P.z_min = z_min; P.z_max = z_max;
    P.zmp_plan = zmp_plan;
    P.tile_counter = nullptr;
    P.num_units = (unsigned)((ncols + kGroupCols - 1) / kGroupCols);
    const int rblk = plan->NP / 8;
    if (plan->N <= kMaxResidentHorizon) {
        switch ((rblk + kRowGroups - 1) / kRowGroups) {
            case 1: return launch_dense<1, 4, false>(plan, P, stream);
            case 2: return launch_dense<2, 4, false>(plan, P, stream);
            case 3: return launch_dense<3, 4, false>(plan, P, stream);
            case 4: return launch_dense<4, 4, false>(plan, P, stream);
            default: return plan->NP <= 152 ? launch_dense<5, 4, false>(plan, P, stream) : launch_dense<5, 2, false>(plan, P, stream);
        }
    }
    const int nrb = (rblk + kRowGroups - 1) / kRowGroups;      // up to 20 row blocks per warp at N = 640
    if (nrb <= 10) return launch_dense<10, 4, true>(plan, P, stream);
    if (nrb <= 20) return launch_dense<20, 4, true>(plan, P, stream);
    set_error("unconstrained preview: horizon %d exceeds the dense kernel's limit 640", plan->N);
    return LIPM_E_UNSUPPORTED;
}

int launch_jerk_from_zmp(lipm_plan* plan, int ncols, const double* state, const double* zmp_plan,
                         double* jerk_plan, cudaStream_t stream) {
    if (ncols <= 0) return LIPM_OK;
    k_jerk_from_zmp<<<ncols, 128, sizeof(double) * 2 * plan->N, stream>>>(plan->N, ncols, plan->d_tinv, plan->d_px,
                                                                           state, zmp_plan, jerk_plan);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

}  // namespace lipm
