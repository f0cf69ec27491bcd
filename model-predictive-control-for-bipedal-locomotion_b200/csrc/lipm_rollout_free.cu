// Unconstrained ("non-strict") mode and the small dense helpers of the single-tick API.
//
// Reference: predict_wieber_axis, non-strict branch (zmp_controller.py:196-199)
//     X = -inv(Pu'Pu + R/Q I) Pu' (Px x - zref);   next = A x + B X[0]
// rebuilt (O(N^2) Python loop) and re-inverted (O(N^3)) on every call.  Only X[0] is consumed, and
//     X[0] = kz . zref_window + kx . x
// with a gain row (kz, kx) that depends on the plan only.  A walk is therefore a sliding
// correlation of the zref series with kz (independent of the state, done 32 ticks at a time by
// the lanes of a warp) followed by the 3-state recursion.  The kernel is bandwidth/latency bound:
// per solve it reads one new (z_min, z_max) sample (16 B) and writes the state (24 B) and a flag.
#include "lipm_internal.cuh"
#include "lipm_device.cuh"

namespace lipm {

namespace {

constexpr int kFreeWarps = 4;   // columns per CTA (one warp per column)

struct FreeParams {
    ColumnBatch cb;
    int N;
    double viol_tol;
    double T, c2, c3, hg;
    const double* gain;   // kz[N], kx[3]
};

__global__ void __launch_bounds__(32 * kFreeWarps) k_free_rollout(const FreeParams P) {
    extern __shared__ double sm[];
    const int N = P.N;
    double* sK = sm;                                  // kz
    double* sZ = sm + N + (threadIdx.x >> 5) * (N + 32);  // this warp's zref samples
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ColumnBatch& B = P.cb;

    for (int j = threadIdx.x; j < N; j += blockDim.x) sK[j] = P.gain[j];
    const double kx0 = P.gain[N], kx1 = P.gain[N + 1], kx2 = P.gain[N + 2];
    __syncthreads();

    const int c = blockIdx.x * kFreeWarps + warp;
    if (c >= B.ncols) return;
    const int axis = c / B.B, b = c - axis * B.B;
    int n = B.n_steps ? B.n_steps[b] : B.n_uniform;
    n = n > B.n_max ? B.n_max : n;
    int ticks = B.ticks_override > 0 ? B.ticks_override : n - 1;
    if (n < 1 || ticks < 0) ticks = 0;
    const long long bz = B.bounds_index ? B.bounds_index[b] : ((B.bounds_batch == 1) ? 0 : b);
    const double* plo = B.z_min + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
    const double* phi = B.z_max + bz * B.bound_scn_stride + axis * B.bound_axis_offset;
    double* hist = axis ? B.hist[1] : B.hist[0];
    if (hist) hist += (long long)b * B.hist_scn_stride;
    uint8_t* pf = B.flags ? B.flags + ((long long)b * B.n_max) * 2 + axis : nullptr;
    const double* init = (axis ? B.init[1] : B.init[0]) + (long long)b * 3;
    double x0 = init[0], x1 = init[1], x2 = init[2];
    const bool push = axis == 1 && B.impulse != nullptr;
    const double imp = push ? B.impulse[b] : 0.0;
    const int ftick = push ? (B.force_tick ? B.force_tick[b] : -1) : -1;

    double viol = 0.0;
    if (B.write_initial && ticks > 0) {
        const double zmp = x0 - P.hg * x2;
        const double below = plo[0] - zmp, above = zmp - phi[0];
        viol = fmax(fmax(below, above), 0.0);
        if (lane == 0) {
            if (hist) { hist[0] = x0; hist[1] = x1; hist[2] = x2; }
            if (pf) pf[0] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                      (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
        }
    }

    const int last = n - 1;
    for (int i0 = 0; i0 < ticks; i0 += 32) {
        // samples i0+bias .. i0+bias+N+30 cover the windows of ticks i0 .. i0+31
        __syncwarp();
        for (int k = lane; k < N + 31; k += 32) {
            int s = i0 + B.bound_index_bias + k;
            s = s > last ? last : s;
            sZ[k] = (__ldg(plo + (long long)s * B.bound_stride) + __ldg(phi + (long long)s * B.bound_stride)) / 2;
        }
        __syncwarp();
        // lane l: state-independent part of the first jerk of tick i0 + l
        double f = 0.0;
        for (int j = 0; j < N; ++j) f = fma(sK[j], sZ[lane + j], f);
        // bounds of the sample each tick produces (window row 0)
        const int my_tick = i0 + lane;
        int s_row0 = my_tick + B.bound_index_bias;
        s_row0 = s_row0 > last ? last : s_row0;
        const double l0 = __ldg(plo + (long long)s_row0 * B.bound_stride);
        const double h0 = __ldg(phi + (long long)s_row0 * B.bound_stride);

        // the recursion is sequential; every lane runs it (warp-uniform) and keeps its own tick
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
        const int steps = (ticks - i0) < 32 ? (ticks - i0) : 32;
        for (int l = 0; l < steps; ++l) {
            const double fl = __shfl_sync(0xffffffffu, f, l);
            const double jerk = fl + fma(kx2, x2, fma(kx1, x1, kx0 * x0));
            const double n0 = fma(P.c3, jerk, fma(P.c2, x2, fma(P.T, x1, x0)));
            double n1 = fma(P.c2, jerk, fma(P.T, x2, x1));
            const double n2 = fma(P.T, jerk, x2);
            if (i0 + l == ftick) n1 -= imp;
            x0 = n0; x1 = n1; x2 = n2;
            if (l == lane) { m0 = n0; m1 = n1; m2 = n2; }
        }
        if (lane < steps) {
            const double zmp = m0 - P.hg * m2;
            const double below = l0 - zmp, above = zmp - h0;
            viol = fmax(viol, fmax(fmax(below, above), 0.0));
            const long long sample = my_tick + 1;
            if (hist) {
                double* dst = hist + (sample - B.hist_index_bias) * 3;
                dst[0] = m0; dst[1] = m1; dst[2] = m2;
            }
            if (pf) pf[sample * 2] = (uint8_t)((below > P.viol_tol ? LIPM_FLAG_VIOL_LO : 0u) |
                                               (above > P.viol_tol ? LIPM_FLAG_VIOL_HI : 0u));
        }
    }
    for (int o = 16; o; o >>= 1) viol = fmax(viol, __shfl_xor_sync(0xffffffffu, viol, o));
    if (lane == 0) {
        if (B.max_violation) B.max_violation[b * 2 + axis] = viol;
        if (B.status) B.status[b * 2 + axis] = 0;
        if (B.iters) B.iters[b * 2 + axis] = 0;
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
    LIPM_CUDA_TRY(cudaMemsetAsync(plan->d_tile_counter, 0, sizeof(unsigned int), stream));
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
    const size_t smem = sizeof(double) * ((size_t)plan->N + (size_t)kFreeWarps * (plan->N + 32));
    if (smem > 48 * 1024)
        LIPM_CUDA_TRY(cudaFuncSetAttribute(k_free_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((cb.ncols + kFreeWarps - 1) / kFreeWarps);
    k_free_rollout<<<grid, 32 * kFreeWarps, smem, stream>>>(P);
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
    P.tile_counter = plan->d_tile_counter;
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
