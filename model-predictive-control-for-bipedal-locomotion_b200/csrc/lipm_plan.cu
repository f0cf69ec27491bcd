// Plan construction: everything that is fixed for one (N, h, g, dt, Q, R, rho) is built on the
// device, once.  Replaces the per-call matrix construction of the reference's
// predict_wieber_axis (zmp_controller.py:162-171) and its per-call inverse (:198).
//
// ZMP coordinates.  With z = Px x + Pu J (the previewed ZMP itself) the per-tick QP
//     min  Q/2 |Px x + Pu J - zref|^2 + R/2 |J|^2     s.t.  zmin <= Px x + Pu J <= zmax
// becomes a box-constrained QP in z with Hessian H = Q I + G, G = R Pu^-T Pu^-1, and linear term
// q = -Q zref - G Px x.  cond(H) ~ 8 (vs ~6e4 for the jerk-coordinate Hessian), the constraint is
// a plain box, and the applied jerk is J0 = (z0 - Px[0] x) / Pu[0,0] because Pu is lower
// triangular.  Pu is Toeplitz, so Pu^-1 is the lower-triangular Toeplitz matrix of the
// first column t, obtained by forward substitution.
#include "lipm_internal.cuh"

#include <cmath>

#include <cuda_fp16.h>

namespace lipm {

namespace {

// ---- model, Px, Pu ---------------------------------------------------------------------------
// The scalar coefficients are evaluated on the host exactly like the Python expressions of the
// reference ((T ** 3) / 6, (T ** 2) / 2, T * h / g, h / g); the per-entry arithmetic below uses
// explicitly rounded operations (no FMA contraction) so that Px and Pu are bit-identical to the
// reference's float64 loop.
struct ModelScalars {
    double T, c3, c2, thg, hg;   // dt, T^3/6, T^2/2, T*h/g, h/g
};

__global__ void k_model(ModelScalars s, int N, double* __restrict__ px, double* __restrict__ pu,
                        double* __restrict__ abc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const double A[9] = {1.0, s.T, s.c2, 0.0, 1.0, s.T, 0.0, 0.0, 1.0};
        for (int k = 0; k < 9; ++k) abc[k] = A[k];
        abc[9] = s.c3; abc[10] = s.c2; abc[11] = s.T;          // B
        abc[12] = 1.0; abc[13] = 0.0; abc[14] = -s.hg;          // C
    }
    if (i >= N) return;
    px[i * 3 + 0] = 1.0;
    px[i * 3 + 1] = __dmul_rn(s.T, (double)(i + 1));
    px[i * 3 + 2] = __dsub_rn(__dmul_rn(s.c2, (double)((long long)(i + 1) * (i + 1))), s.hg);
    for (int j = 0; j < N; ++j) {
        const long long d = i - j;
        pu[(size_t)i * N + j] =
            (d >= 0) ? __dsub_rn(__dmul_rn(s.c3, (double)(1 + 3 * d + 3 * d * d)), s.thg) : 0.0;
    }
}

// ---- first column of Pu^-1 -------------------------------------------------------------------
// t0 = 1/p0,  t_k = -(sum_{j<k} p_{k-j} t_j) / p0.  One CTA; the dot products are block reductions.
__global__ void k_toeplitz_inverse(int N, const double* __restrict__ pu, double* __restrict__ tinv) {
    extern __shared__ double sm[];
    double* p = sm;            // first column of Pu
    double* t = sm + N;
    __shared__ double partial[32];
    for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = pu[(size_t)i * N];
    __syncthreads();
    if (threadIdx.x == 0) t[0] = 1.0 / p[0];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = 1; k < N; ++k) {
        double acc = 0.0;
        for (int j = threadIdx.x; j < k; j += blockDim.x) acc = fma(p[k - j], t[j], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) partial[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += partial[w];
            t[k] = -s / p[0];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) tinv[i] = t[i];
}

// ---- G = R T^T T with T = lower-triangular Toeplitz(t) ----------------------------------------
__global__ void k_gram(int N, double R, const double* __restrict__ tinv, double* __restrict__ G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N || i >= N) return;
    const int k0 = max(i, j);
    double acc = 0.0;
    for (int k = k0; k < N; ++k) acc = fma(tinv[k - i], tinv[k - j], acc);
    G[(size_t)i * N + j] = R * acc;
}

__global__ void k_gpx(int N, int NP, const double* __restrict__ G, const double* __restrict__ px,
                      double* __restrict__ gpx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (i < N) {
        for (int j = 0; j < N; ++j) {
            const double gij = G[(size_t)i * N + j];
            a0 = fma(gij, px[j * 3 + 0], a0);
            a1 = fma(gij, px[j * 3 + 1], a1);
            a2 = fma(gij, px[j * 3 + 2], a2);
        }
    }
    gpx[i * 3 + 0] = a0; gpx[i * 3 + 1] = a1; gpx[i * 3 + 2] = a2;
}

// ---- W = G + diag  ->  Cholesky (lower, in place)  ->  W^-1 ------------------------------------
__global__ void k_shifted_copy(int N, double diag, const double* __restrict__ G, double* __restrict__ W) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * N) return;
    const int i = (int)(idx / N), j = (int)(idx % N);
    W[idx] = G[idx] + (i == j ? diag : 0.0);
}

// Right-looking Cholesky by one CTA.  The matrices are tiny (N <= 600) and this runs once per plan.
__global__ void k_cholesky(int N, double* __restrict__ W, int* __restrict__ info) {
    __shared__ double pivot;
    for (int k = 0; k < N; ++k) {
        if (threadIdx.x == 0) {
            const double d = W[(size_t)k * N + k];
            if (!(d > 0.0) || !isfinite(d)) { atomicExch(info, k + 1); pivot = 1.0; }
            else pivot = sqrt(d);
            W[(size_t)k * N + k] = pivot;
        }
        __syncthreads();
        const double piv = pivot;
        for (int i = k + 1 + threadIdx.x; i < N; i += blockDim.x) W[(size_t)i * N + k] /= piv;
        __syncthreads();
        // trailing update of the lower triangle: W[i][j] -= L[i][k] L[j][k], k < j <= i
        const int m = N - k - 1;
        for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
            const int i = k + 1 + idx / m, j = k + 1 + idx % m;
            if (j <= i) W[(size_t)i * N + j] = fma(-W[(size_t)i * N + k], W[(size_t)j * N + k], W[(size_t)i * N + j]);
        }
        __syncthreads();
    }
}

// Column j of W^-1 from the factor: L y = e_j, then L^T m = y.  One thread per column; the
// column lives in out[.][j] so that neighbouring threads touch neighbouring addresses.
__global__ void k_cholesky_inverse(int N, const double* __restrict__ L, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    for (int i = 0; i < j; ++i) out[(size_t)i * N + j] = 0.0;
    for (int i = j; i < N; ++i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) s = fma(-L[(size_t)i * N + k], out[(size_t)k * N + j], s);
        out[(size_t)i * N + j] = s / L[(size_t)i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
        double s = out[(size_t)i * N + j];
        for (int k = i + 1; k < N; ++k) s = fma(-L[(size_t)k * N + i], out[(size_t)k * N + j], s);
        out[(size_t)i * N + j] = s / L[(size_t)i * N + i];
    }
}

// Exact symmetrisation (the two triangles agree to rounding only).
__global__ void k_symmetrise(int N, double* __restrict__ M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (i >= N || j >= N || j <= i) return;
    const double v = 0.5 * (M[(size_t)i * N + j] + M[(size_t)j * N + i]);
    M[(size_t)i * N + j] = v;
    M[(size_t)j * N + i] = v;
}

// ADMM matrix in DMMA.8x8x4 A-fragment order: tile (I, kb) is 8 rows x 4 columns, lane l holds
// element (row l/4, column l%4); tiles are stored [I][kb][lane] so a warp's fragment load is one
// contiguous 256-byte read.  Rows / columns beyond N are zero.
__global__ void k_pack_fragments(int N, int NP, const double* __restrict__ M, double* __restrict__ frag) {
    const int KB = NP / 4;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)(NP / 8) * KB * 32;
    if (idx >= total) return;
    const int lane = (int)(idx & 31);
    const size_t tile = idx >> 5;
    const int kb = (int)(tile % KB), I = (int)(tile / KB);
    const int row = I * 8 + (lane >> 2), col = kb * 4 + (lane & 3);
    frag[idx] = (row < N && col < N) ? M[(size_t)row * N + col] : 0.0;
}

// First-jerk gain row of the unconstrained mode:  z* = H^-1 (Q zref + G Px x),
// u0 = (z*_0 - Px[0] x) / Pu00  =  kz . zref + kx . x.
__global__ void k_gain(int N, double Q, double pu00, const double* __restrict__ hinv,
                       const double* __restrict__ gpx, const double* __restrict__ px,
                       double* __restrict__ gain) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) gain[j] = Q * hinv[j] / pu00;
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        double s = 0.0;
        for (int j = 0; j < N; ++j) s = fma(hinv[j], gpx[j * 3 + c], s);
        gain[N + c] = (s - px[c]) / pu00;
    }
}

// B operand of the reduced-precision kernel: row n (output row), contraction index k:
//   k < N            -> M[n][k]
//   k = 152 + c      -> (M G Px)[n][c]   (multiplies the axis state carried in the K padding)
// split into FP16 hi + lo (hi = rn(v), lo = rn(v - hi)) and laid out as the tensor core reads it.
__global__ void k_pack_f16(int N, const double* __restrict__ M, const double* __restrict__ gpx,
                           unsigned char* __restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kF16Tile * kF16Tile) return;
    const int n = idx / kF16Tile, k = idx % kF16Tile;
    double v = 0.0;
    if (n < N && k < N) v = M[(size_t)n * N + k];
    else if (n < N && k >= kF16StateSlot && k < kF16StateSlot + 3) {
        const int c = k - kF16StateSlot;
        for (int j = 0; j < N; ++j) v = fma(M[(size_t)n * N + j], gpx[j * 3 + c], v);
    }
    const __half hi = __double2half(v);
    const __half lo = __double2half(v - (double)__half2float(hi));
    const int off = f16_operand_offset(n, k, kF16Tile);
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + kF16OperandBytesB + off) = lo;
}

// The same matrix for the streamed kernel (lipm_rollout_strict_f16s.cu): output rows cut into 160-row slabs,
// contraction length NK (multiple of 16), state slots at k = N .. N+2; image layout [slab][k / 8][row][k % 8].
__global__ void k_pack_f16s(int N, int NK, int slabs, const double* __restrict__ M, const double* __restrict__ gpx,
                            unsigned char* __restrict__ img, size_t image_bytes) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)slabs * kF16Tile * NK;
    if (idx >= total) return;
    const int k = (int)(idx % NK);
    const int row = (int)(idx / NK);               // 0 .. slabs * 160 - 1
    const int slab = row / kF16Tile, r = row % kF16Tile;
    double v = 0.0;
    if (row < N && k < N) v = M[(size_t)row * N + k];
    else if (row < N && k < N + 3) {
        const int c = k - N;
        for (int j = 0; j < N; ++j) v = fma(M[(size_t)row * N + j], gpx[j * 3 + c], v);
    }
    const __half hi = __double2half(v);
    const __half lo = __double2half(v - (double)__half2float(hi));
    const size_t off = (((size_t)slab * (NK / 8) + (k >> 3)) * kF16Tile + r) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(img + off) = hi;
    *reinterpret_cast<__half*>(img + image_bytes + off) = lo;
}

// Largest eigenvalue of the symmetric positive definite G by power iteration (one CTA; N <= 4096).
__global__ void k_power_iteration(int N, const double* __restrict__ G, int iterations, double* __restrict__ scratch,
                                  double* __restrict__ lambda_out) {
    __shared__ double red[32];
    __shared__ double norm;
    double* x = scratch;
    double* y = scratch + N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] = 1.0 + 1e-3 * (i % 7);
    __syncthreads();
    double lambda = 0.0;
    for (int it = 0; it < iterations; ++it) {
        double part = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double acc = 0.0;
            for (int j = 0; j < N; ++j) acc = fma(G[(size_t)i * N + j], x[j], acc);
            y[i] = acc;
            part = fma(acc, acc, part);
        }
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            norm = sqrt(s);
        }
        __syncthreads();
        lambda = norm;                                 // |G x| with |x| = 1 (after the first pass)
        for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] = y[i] / norm;
        __syncthreads();
    }
    if (threadIdx.x == 0) *lambda_out = lambda;
}

int invert_shifted(lipm_plan* plan, double diag, double* d_out) {
    const int N = plan->N;
    const size_t nn = (size_t)N * N;
    k_shifted_copy<<<(unsigned)((nn + 255) / 256), 256>>>(N, diag, plan->d_g, plan->d_work);
    k_cholesky<<<1, 1024>>>(N, plan->d_work, plan->d_info);
    k_cholesky_inverse<<<(N + 63) / 64, 64>>>(N, plan->d_work, d_out);
    k_symmetrise<<<dim3((N + 127) / 128, N), 128>>>(N, d_out);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

}  // namespace

int build_plan_matrices(lipm_plan* plan) {
    const int N = plan->N, NP = plan->NP;
    const lipm_plan_desc_t& d = plan->desc;
    const size_t nn = (size_t)N * N;

    ModelScalars s;
    s.T = d.dt;
    s.c3 = std::pow(d.dt, 3) / 6.0;
    s.c2 = std::pow(d.dt, 2) / 2.0;
    s.thg = d.dt * d.h / d.g;
    s.hg = d.h / d.g;
    plan->T = d.dt;
    plan->hg = s.hg;
    plan->pu00 = s.c3 - s.thg;

    LIPM_CUDA_TRY(cudaMalloc(&plan->d_px, sizeof(double) * N * 3));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_pu, sizeof(double) * nn));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_abc, sizeof(double) * 15));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_tinv, sizeof(double) * N));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_g, sizeof(double) * nn));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_gpx, sizeof(double) * NP * 3));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_work, sizeof(double) * nn));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_hinv, sizeof(double) * nn));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_minv, sizeof(double) * nn));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_mfrag, sizeof(double) * (size_t)(NP / 8) * (NP / 4) * 32));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_gain, sizeof(double) * (N + 3)));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_info, sizeof(int) * 4));
    LIPM_CUDA_TRY(cudaMemset(plan->d_info, 0, sizeof(int) * 4));
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_stats, 4 * sizeof(unsigned long long)));
    LIPM_CUDA_TRY(cudaMemset(plan->d_stats, 0, 4 * sizeof(unsigned long long)));

    k_model<<<(N + 127) / 128, 128>>>(s, N, plan->d_px, plan->d_pu, plan->d_abc);
    k_toeplitz_inverse<<<1, 256, sizeof(double) * 2 * N>>>(N, plan->d_pu, plan->d_tinv);
    k_gram<<<dim3((N + 127) / 128, N), 128>>>(N, d.R, plan->d_tinv, plan->d_g);
    k_gpx<<<(NP + 127) / 128, 128>>>(N, NP, plan->d_g, plan->d_px, plan->d_gpx);
    LIPM_CUDA_TRY(cudaGetLastError());

    if (!(plan->desc.rho > 0)) {
        // ADMM penalty from the spectrum of H = Q I + G: rho = sqrt(lambda_min lambda_max) = sqrt(Q (Q + lambda_max(G))).
        // For N <= 160 this is 2.9 and the tuned 2.0 is kept (lipm_api.cu); cond(H) grows to 28 (N=300) and 110
        // (N=600), where rho = 2 needs 1.6x / 3x the iterations of this rule.
        double* d_scratch = nullptr;
        LIPM_CUDA_TRY(cudaMalloc(&d_scratch, sizeof(double) * (2 * (size_t)N + 1)));
        k_power_iteration<<<1, 1024>>>(N, plan->d_g, 60, d_scratch, d_scratch + 2 * N);
        double lambda_max = 0.0;
        LIPM_CUDA_TRY(cudaMemcpy(&lambda_max, d_scratch + 2 * N, sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d_scratch);
        if (!(lambda_max > 0.0) || !std::isfinite(lambda_max)) { set_error("power iteration on G failed"); return LIPM_E_NUMERIC; }
        plan->desc.rho = std::sqrt(d.Q * (d.Q + lambda_max));
    }
    int rc = invert_shifted(plan, d.Q, plan->d_hinv);
    if (rc != LIPM_OK) return rc;
    rc = invert_shifted(plan, d.Q + plan->desc.rho, plan->d_minv);
    if (rc != LIPM_OK) return rc;

    const size_t nfrag = (size_t)(NP / 8) * (NP / 4) * 32;
    k_pack_fragments<<<(unsigned)((nfrag + 255) / 256), 256>>>(N, NP, plan->d_minv, plan->d_mfrag);
    LIPM_CUDA_TRY(cudaMalloc(&plan->d_hfrag, sizeof(double) * nfrag));
    k_pack_fragments<<<(unsigned)((nfrag + 255) / 256), 256>>>(N, NP, plan->d_hinv, plan->d_hfrag);
    k_gain<<<1, 256>>>(N, d.Q, plan->pu00, plan->d_hinv, plan->d_gpx, plan->d_px, plan->d_gain);
    if (d.precision == LIPM_PRECISION_F16X3 && N <= kMaxReducedHorizon) {
        LIPM_CUDA_TRY(cudaMalloc(&plan->d_bf16, 2 * kF16OperandBytesB));
        k_pack_f16<<<(kF16Tile * kF16Tile + 255) / 256, 256>>>(N, plan->d_minv, plan->d_gpx, plan->d_bf16);
    } else if (d.precision == LIPM_PRECISION_F16X3) {
        plan->f16s_nk = (N + 3 + 15) & ~15;
        plan->f16s_slabs = (plan->f16s_nk + kF16Tile - 1) / kF16Tile;
        const size_t image = (size_t)plan->f16s_slabs * (plan->f16s_nk / 8) * kF16Tile * 16;
        LIPM_CUDA_TRY(cudaMalloc(&plan->d_f16s, 2 * image));
        const long long total = (long long)plan->f16s_slabs * kF16Tile * plan->f16s_nk;
        k_pack_f16s<<<(unsigned)((total + 255) / 256), 256>>>(N, plan->f16s_nk, plan->f16s_slabs, plan->d_minv, plan->d_gpx,
                                                              plan->d_f16s, image);
    }
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(cudaDeviceSynchronize());

    rc = build_riccati_tables(plan);
    if (rc != LIPM_OK) return rc;
    rc = build_rowpar_tables(plan);
    if (rc != LIPM_OK) return rc;
    LIPM_CUDA_TRY(cudaDeviceSynchronize());

    int info[4];
    LIPM_CUDA_TRY(cudaMemcpy(info, plan->d_info, sizeof(info), cudaMemcpyDeviceToHost));
    if (info[0] != 0) {
        set_error("Cholesky of the ZMP-coordinate Hessian broke down at pivot %d (N=%d, Q=%g, R=%g)",
                  info[0] - 1, N, d.Q, d.R);
        return LIPM_E_NUMERIC;
    }
    return LIPM_OK;
}

}  // namespace lipm
