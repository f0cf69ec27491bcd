// Herdt free-footstep MPC (reference zmp_controller.py:435-826), batched over scenarios that share a
// support timeline.  The reference builds, per tick, one joint x/y QP over u = [Jx, Fx, Jy, Fy]
// (jerks + m future footsteps, :533-826) and hands it to cvxpy -> OSQP.
//
// Coordinates.  Per axis xi = (e, F) with e = z - U F - U_c fc (previewed ZMP relative to the foot that
// supports it) and F the future footsteps.  Then
//   * the ZMP-around-foot rows (:661-713) are a plain box |e_i| <= half on the rows that are not STANDING,
//     the standing-hull rows (:715-771) a box on e_i as well (rows supported by the current foot; the one
//     exception is handled by "absolute" rows, see k_herdt_assemble), and the footstep polytope (:773-790) a 2-D projection of the pair (F1x, F1y): the only
//     place where the two axes meet;
//   * the cost 1/2 alpha |J|^2 + 1/2 beta |vel - vref|^2 + 1/2 gamma |e|^2 has the Hessian
//     P = S'(alpha T'T + beta V'V) S + gamma diag(I, 0),  S = [I U], T = Pzu^-1, V = Pvu T: it depends on the
//     timeline and the tick only -- ONE matrix for both axes and for every scenario of the group;
//   * J0 = (e_0 + fc - Pzx[0] x) / Pzu[0,0].
// The QP is solved by the same fixed-rho ADMM in t-form as the strict Wieber kernel (lipm_rollout_strict.cu):
// v = M (rho (2z - t) - q), M = (P + rho I)^-1 in DMMA fragment order resident in shared memory, epilogue =
// relaxation + projection (box, or polygon for the footstep pair) + residual, FP64 tensor pipe.
// The x and y columns of a scenario are adjacent, so the thread that owns the F1 row holds both components
// of the pair in its accumulator fragment and projects them on the polygon locally.
//
// Everything is built on the device: T, V (Toeplitz first columns), W = alpha T'T + beta V'V once per plan;
// P of every tick of a timeline from the phase layout; batched Cholesky inverses; fragment packing; the
// linear term q = S'(alpha T'(T d1) + beta V'(V d1 + d2)) by Toeplitz products per column.
#include "lipm_internal.cuh"
#include "lipm_device.cuh"

#include <cmath>
#include <cstring>
#include <new>

namespace lipm {
namespace {

constexpr int kMaxPolyVerts = 16;
constexpr double kFree = 1e30;     // "no bound"

// ------------------------------------------------------------------------------------------ plan kernels
// first columns of the lower-triangular Toeplitz matrices Pzu, Pvu (zmp_controller.py:576-579)
__global__ void k_herdt_toeplitz(int N, double T, double hg, double* __restrict__ pzu, double* __restrict__ pvu) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= N) return;
    const double c3 = (T * T * T) / 6.0, c2 = (T * T) / 2.0;
    pzu[d] = c3 * (double)(1 + 3 * d + 3 * d * d) - T * hg;
    pvu[d] = c2 * (double)(2 * d + 1);
}

// t = first column of Pzu^-1 (forward substitution), v = first column of V = Pvu Pzu^-1 (convolution)
__global__ void k_herdt_tv(int N, const double* __restrict__ pzu, const double* __restrict__ pvu,
                           double* __restrict__ tcol, double* __restrict__ vcol) {
    extern __shared__ double sm[];
    double* p = sm;
    double* t = sm + N;
    __shared__ double partial[32];
    for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = pzu[i];
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
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        tcol[i] = t[i];
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc = fma(pvu[i - j], t[j], acc);
        vcol[i] = acc;
    }
}

// W = alpha T'T + beta V'V  (T, V lower-triangular Toeplitz)
__global__ void k_herdt_w(int N, double alpha, double beta, const double* __restrict__ t, const double* __restrict__ v,
                          double* __restrict__ W) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N || i >= N) return;
    double a = 0.0, b = 0.0;
    for (int k = max(i, j); k < N; ++k) {
        a = fma(t[k - i], t[k - j], a);
        b = fma(v[k - i], v[k - j], b);
    }
    W[(size_t)i * N + j] = alpha * a + beta * b;
}

// ------------------------------------------------------------------------------------------ timeline kernels
// P of one tick: n = N + m_pad variables (e rows then footsteps; unused footstep slots get a unit diagonal)
// A row is "absolute" when it is a STANDING sample constrained to the current hull (kind 2) although a FUTURE
// footstep supports it (the last SINGLE_SUPPORT tick before the final stance): its variable is the ZMP itself
// (z_i = e_i), so that the hull stays a box; the tracking term gamma (e_i - F_k)^2 then couples e_i and F_k.
__device__ __forceinline__ bool absolute_row(const int8_t* phase, const int8_t* kind, int j) {
    return kind[j] == 2 && phase[j] >= 0;
}

__global__ void k_herdt_assemble(int N, int m_pad, double gamma, const double* __restrict__ W,
                                 const int8_t* __restrict__ phase_all, const int8_t* __restrict__ kind_all,
                                 double* __restrict__ P_all) {
    extern __shared__ double C[];                     // [N][m_pad]: C[i][k] = sum_{j in phase k, relative} W[i][j]
    __shared__ int used[16];
    const int tick = blockIdx.x;
    const int n = N + m_pad;
    const int8_t* phase = phase_all + (size_t)tick * N;
    const int8_t* kind = kind_all + (size_t)tick * N;
    double* P = P_all + (size_t)tick * n * n;
    if (threadIdx.x < 16) used[threadIdx.x] = 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * m_pad; idx += blockDim.x) {
        const int i = idx / m_pad, k = idx % m_pad;
        double acc = 0.0;
        for (int j = 0; j < N; ++j)
            if (phase[j] == k && !absolute_row(phase, kind, j)) acc += W[(size_t)i * N + j];
        C[idx] = acc;
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x)
        if (phase[j] >= 0) used[phase[j]] = 1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx % n;
        double val;
        if (r < N && c < N) val = W[(size_t)r * N + c] + (r == c ? gamma : 0.0);
        else if (r < N || c < N) {
            const int row = r < N ? r : c, k = (r < N ? c : r) - N;
            val = C[row * m_pad + k];
            if (absolute_row(phase, kind, row) && phase[row] == k) val -= gamma;
        } else {
            const int k = r - N, k2 = c - N;
            double acc = 0.0;
            for (int i = 0; i < N; ++i) {
                if (phase[i] != k) continue;
                if (absolute_row(phase, kind, i)) acc += (k == k2) ? gamma : 0.0;
                else acc += C[i * m_pad + k2];
            }
            val = acc;
            if (k == k2 && !used[k]) val = 1.0;       // footstep variable that appears nowhere: pinned to 0
        }
        P[idx] = val;
    }
}

__global__ void k_shift_diag_batched(int n, double diag, double* __restrict__ P_all) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P_all[(size_t)blockIdx.y * n * n + (size_t)i * n + i] += diag;
}

// Right-looking Cholesky, one CTA per matrix (in place, lower triangle).
__global__ void k_cholesky_batched(int n, double* __restrict__ W_all, int* __restrict__ info) {
    double* W = W_all + (size_t)blockIdx.x * n * n;
    __shared__ double pivot;
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) {
            const double d = W[(size_t)k * n + k];
            if (!(d > 0.0) || !isfinite(d)) { atomicExch(info, k + 1); pivot = 1.0; }
            else pivot = sqrt(d);
            W[(size_t)k * n + k] = pivot;
        }
        __syncthreads();
        const double piv = pivot;
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) W[(size_t)i * n + k] /= piv;
        __syncthreads();
        const int mrem = n - k - 1;
        for (int idx = threadIdx.x; idx < mrem * mrem; idx += blockDim.x) {
            const int i = k + 1 + idx / mrem, j = k + 1 + idx % mrem;
            if (j <= i) W[(size_t)i * n + j] = fma(-W[(size_t)i * n + k], W[(size_t)j * n + k], W[(size_t)i * n + j]);
        }
        __syncthreads();
    }
}

// Column j of the inverse from the factor (L y = e_j, L' x = y); grid.y = matrix.
__global__ void k_cholesky_inverse_batched(int n, const double* __restrict__ L_all, double* __restrict__ out_all) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double* L = L_all + (size_t)blockIdx.y * n * n;
    double* out = out_all + (size_t)blockIdx.y * n * n;
    for (int i = 0; i < j; ++i) out[(size_t)i * n + j] = 0.0;
    for (int i = j; i < n; ++i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) s = fma(-L[(size_t)i * n + k], out[(size_t)k * n + j], s);
        out[(size_t)i * n + j] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = out[(size_t)i * n + j];
        for (int k = i + 1; k < n; ++k) s = fma(-L[(size_t)k * n + i], out[(size_t)k * n + j], s);
        out[(size_t)i * n + j] = s / L[(size_t)i * n + i];
    }
}

// symmetrise + pack into DMMA.8x8x4 A-fragment order ([NP/8][NP/4][32]); grid.y = matrix
__global__ void k_pack_fragments_batched(int n, int NP, const double* __restrict__ M_all, double* __restrict__ frag_all) {
    const int KB = NP / 4;
    const size_t total = (size_t)(NP / 8) * KB * 32;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const double* M = M_all + (size_t)blockIdx.y * n * n;
    const int lane = (int)(idx & 31);
    const size_t tile = idx >> 5;
    const int kb = (int)(tile % KB), I = (int)(tile / KB);
    const int row = I * 8 + (lane >> 2), col = kb * 4 + (lane & 3);
    double v = 0.0;
    if (row < n && col < n) v = 0.5 * (M[(size_t)row * n + col] + M[(size_t)col * n + row]);
    frag_all[(size_t)blockIdx.y * total + idx] = v;
}

// ------------------------------------------------------------------------------------------ per tick
struct TickParams {
    int N, m_pad, NP, m;          // horizon, footstep slots, padded size, footsteps of this tick
    int B;                        // scenarios (columns = 2B: x, y adjacent)
    int tick, n_ext;              // window rows read vref samples tick+1+j (clamped to n_ext-1)
    int vref_batch;               // 1 or B
    int foot_side;                // 0 left, 1 right (current support foot)
    double T, hg, alpha, beta;
    double half[2];               // 0.5 foot_length, 0.5 foot_width
    double foot_spread;
    const double* spread;         // [B] per-scenario foot spread or nullptr (-> foot_spread)
    const int8_t* phase;          // [N] of this tick
    const int8_t* kind;           // [N] 0 free, 1 foot box, 2 standing hull
    const double *tcol, *vcol;    // Toeplitz first columns of T, V
    const double *x, *y;          // [B,3]
    const double* fc;             // [B,2]
    const double* vref;           // [vref_batch, n_ext, 2]
    double *q, *lo, *hi, *t0;     // [2B, NP]
    const double* t_prev;         // [2B, NP] final ADMM iterate t of the previous tick, or nullptr (cold start)
    int shift_f;                  // the previous tick's first future footstep became the current phase
    int m_prev;                   // footsteps of the previous tick
};

// linear term, bounds and start point of every column (one CTA per column)
__global__ void __launch_bounds__(128) k_herdt_prepare(const TickParams P) {
    extern __shared__ double sm[];
    const int N = P.N;
    double* d1 = sm;            // U_c fc - Pzx x
    double* a = sm + N;         // T d1
    double* bb = sm + 2 * N;    // V d1 + d2
    double* qz = sm + 3 * N;
    double* st = sm + 4 * N;    // t column
    double* sv = sm + 5 * N;    // v column
    const int c = blockIdx.x, b = c >> 1, axis = c & 1;
    const double* state = (axis ? P.y : P.x) + (size_t)b * 3;
    const double x0 = state[0], x1 = state[1], x2 = state[2];
    const double fc = P.fc[(size_t)b * 2 + axis];
    const double* vref = P.vref + (P.vref_batch == 1 ? 0 : (size_t)b * P.n_ext * 2) + axis;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double tj = (double)(j + 1) * P.T;
        const double zx = x0 + tj * x1 + (0.5 * tj * tj - P.hg) * x2;              // Pzx[j] . x
        d1[j] = (P.phase[j] == -1 ? fc : 0.0) - zx;
        int smp = P.tick + 1 + j;
        smp = smp > P.n_ext - 1 ? P.n_ext - 1 : smp;
        bb[j] = x1 + tj * x2 - vref[(size_t)smp * 2];                                // d2 = Pvs x - vref
        st[j] = P.tcol[j];
        sv[j] = P.vcol[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j <= i; ++j) {
            s1 = fma(st[i - j], d1[j], s1);
            s2 = fma(sv[i - j], d1[j], s2);
        }
        a[i] = s1;
        qz[i] = s2;     // V d1 (bb is still being read by nobody else; combine after the barrier)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) bb[i] += qz[i];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int k = i; k < N; ++k) {
            s1 = fma(st[k - i], a[k], s1);
            s2 = fma(sv[k - i], bb[k], s2);
        }
        qz[i] = P.alpha * s1 + P.beta * s2;
    }
    __syncthreads();
    const size_t base = (size_t)c * P.NP;
    // standing hull (zmp_controller.py:715-750): x: current foot +- length/2; y: both feet +- width/2
    double box_lo, box_hi;
    if (axis == 0) {
        box_lo = fc - P.half[0];
        box_hi = fc + P.half[0];
    } else {
        const double sp = P.spread ? P.spread[b] : P.foot_spread;
        const double other = P.foot_side == 0 ? fc - 2.0 * sp : fc + 2.0 * sp;
        box_lo = fmin(fc, other) - P.half[1];
        box_hi = fmax(fc, other) + P.half[1];
    }
    for (int j = threadIdx.x; j < P.NP; j += blockDim.x) {
        double q = 0.0, lo = 0.0, hi = 0.0, t0 = 0.0;
        if (j < N) {
            q = qz[j];
            const int kind = P.kind[j];
            if (kind == 1) { lo = -P.half[axis]; hi = P.half[axis]; }
            else if (kind == 2) { const double ref = P.phase[j] == -1 ? fc : 0.0; lo = box_lo - ref; hi = box_hi - ref; }
            else { lo = -kFree; hi = kFree; }
            t0 = fmin(fmax(0.0, lo), hi);
            // warm start: the window slid by one sample and e is relative to the supporting foot, so the previous
            // iterate (primal + scaled dual) of row j+1 is the natural start of row j; hull rows start cold
            if (P.t_prev && kind != 2) t0 = j + 1 < N ? P.t_prev[base + j + 1] : fmin(fmax(P.t_prev[base + j], lo), hi);
        } else if (j < N + P.m) {
            const int k = j - N;
            double s = 0.0;
            for (int i = 0; i < N; ++i)
                if (P.phase[i] == k && !absolute_row(P.phase, P.kind, i)) s += qz[i];
            q = s;
            lo = -kFree; hi = kFree;
            t0 = fc;
            if (P.t_prev && P.m_prev > 0) {
                int kp = k + P.shift_f;
                kp = kp > P.m_prev - 1 ? P.m_prev - 1 : kp;      // a footstep that just entered the window starts at the last known one
                t0 = P.t_prev[base + N + kp];
            }
        }
        P.q[base + j] = q; P.lo[base + j] = lo; P.hi[base + j] = hi; P.t0[base + j] = t0;
    }
}

// --- the QP kernel -------------------------------------------------------------------------------------
struct QpParams {
    int ncols, NP, n;             // columns (even: x/y pairs), padded size, variables
    int f1_row;                   // row of the first footstep (N) or -1
    int max_iter;
    double rho, alpha, eps, stop_factor;
    double accept_eps;            // a pair that stalls / hits max_iter within this (looser) tolerance is still applied; < 0: never
    int stall_window;             // iterations between two stall checks
    const double* mfrag;
    const double *q, *lo, *hi, *t0;   // [ncols, NP]
    double* sol;                  // [ncols, NP] projected solution
    double* t_out;                // [ncols, NP] final iterate (warm start of the next tick) or nullptr
    int32_t *status, *iters;      // [ncols]
    const double* poly_shift;     // [ncols/2, 2]: the polygon is given relative to this point (current foot)
    int n_verts;
    double verts[kMaxPolyVerts][2];   // counter-clockwise
    unsigned int* tile_counter;
    unsigned int num_units;
};

struct QpGroup {
    uint32_t Res[2][kRowGroups][kGroupCols];
    int Iter[kGroupCols], NotConv[kGroupCols];
    int Unit;
};

// Euclidean projection of (px, py) on the convex polygon (vertices counter-clockwise, relative coordinates)
__device__ __forceinline__ void project_polygon(const QpParams& P, double px, double py, double& ox, double& oy) {
    bool inside = true;
    double best = 1e300, bx = px, by = py;
    for (int i = 0; i < P.n_verts; ++i) {
        const int i2 = (i + 1 == P.n_verts) ? 0 : i + 1;
        const double ax = P.verts[i][0], ay = P.verts[i][1];
        const double ex = P.verts[i2][0] - ax, ey = P.verts[i2][1] - ay;
        const double wx = px - ax, wy = py - ay;
        if (ex * wy - ey * wx < 0.0) inside = false;
        double s = (wx * ex + wy * ey) / (ex * ex + ey * ey);
        s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
        const double cx = ax + s * ex, cy = ay + s * ey;
        const double d = (px - cx) * (px - cx) + (py - cy) * (py - cy);
        if (d < best) { best = d; bx = cx; by = cy; }
    }
    ox = inside ? px : bx;
    oy = inside ? py : by;
}

template <int NRB, int NG>
__global__ void __launch_bounds__(32 * kRowGroups * NG, 1) k_qp_tick(const QpParams P) {
    constexpr int kGroupThreads = 32 * kRowGroups;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NP = P.NP;
    const int KB = NP / 4, RBLK = NP / 8, KP = NP + 4;
    const int aStride = KB * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / kRowGroups, gtid = tid & (kGroupThreads - 1);
    const int rg = (warp + grp) & (kRowGroups - 1);
    const int g = lane >> 2, t = lane & 3;
    const int bar_id = 1 + grp;

    unsigned long long* sBar = reinterpret_cast<unsigned long long*>(smem_raw);
    constexpr size_t kFixedStride = (sizeof(QpGroup) + 15) & ~(size_t)15;
    QpGroup& S = *reinterpret_cast<QpGroup*>(smem_raw + 16 + grp * kFixedStride);
    double* sR = reinterpret_cast<double*>(smem_raw + 16 + NG * kFixedStride) + grp * (kGroupCols * KP);
    double* sM = reinterpret_cast<double*>(smem_raw + 16 + NG * kFixedStride) + NG * (kGroupCols * KP);

    if (tid == 0) mbar_init(sBar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t m_bytes = (uint32_t)((size_t)RBLK * KB * 32 * sizeof(double));
        mbar_expect_tx(sBar, m_bytes);
        const uint32_t chunk = 32768;
        for (uint32_t off = 0; off < m_bytes; off += chunk) {
            const uint32_t nbytes = (m_bytes - off < chunk) ? (m_bytes - off) : chunk;
            bulk_load(reinterpret_cast<unsigned char*>(sM) + off, reinterpret_cast<const unsigned char*>(P.mfrag) + off, nbytes, sBar);
        }
    }
    mbar_wait(sBar, 0);

    int nrb = RBLK - rg * NRB;
    nrb = nrb < 0 ? 0 : (nrb > NRB ? NRB : nrb);
    const int row0 = rg * NRB * 8 + g;
    const int lc0 = 2 * t;                                  // columns lc0 (x axis), lc0 + 1 (y axis) of one scenario
    const double* aBase = sM + (size_t)(rg * NRB) * aStride + lane;
    const double* bBase = sR + g * KP + t;
    double* rOut0 = sR + lc0 * KP + row0;
    const double two_rho = 2.0 * P.rho;
    // which of this thread's elements is the footstep pair
    int f1_r = -1;
    if (P.f1_row >= 0) {
        const int rel = P.f1_row - row0;
        if (rel >= 0 && (rel & 7) == 0 && (rel >> 3) < nrb) f1_r = rel >> 3;
    }

    double tv[NRB][2], lo[NRB][2], hi[NRB][2], q[NRB][2];

    for (;;) {
        group_sync<kGroupThreads>(bar_id);
        if (gtid == 0) S.Unit = (int)atomicAdd(P.tile_counter, 1u);
        group_sync<kGroupThreads>(bar_id);
        const unsigned int unit = (unsigned int)S.Unit;
        if (unit >= P.num_units) break;
        const int col0 = (int)unit * kGroupCols;
        if (gtid < kGroupCols) { S.Iter[gtid] = 0; S.NotConv[gtid] = 0; }

        bool live[2];
        uint32_t qmag[2] = {0u, 0u};
        double zf[2] = {0.0, 0.0};                           // projected footstep pair (this thread's, if any)
        double shift[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = col0 + lc0 + e;
            live[e] = c < P.ncols;
            const size_t base = (size_t)(live[e] ? c : 0) * NP;
#pragma unroll
            for (int r = 0; r < NRB; ++r) {
                const int j = row0 + r * 8;
                const bool ok = live[e] && r < nrb;
                tv[r][e] = ok ? P.t0[base + j] : 0.0;
                lo[r][e] = ok ? P.lo[base + j] : 0.0;
                hi[r][e] = ok ? P.hi[base + j] : 0.0;
                q[r][e] = ok ? P.q[base + j] : 0.0;
                qmag[e] = max(qmag[e], mag_hi(q[r][e]));
            }
        }
        if (f1_r >= 0 && live[0]) {
            const int pair = (col0 + lc0) >> 1;
            shift[0] = P.poly_shift[(size_t)pair * 2];
            shift[1] = P.poly_shift[(size_t)pair * 2 + 1];
        }
        // first right-hand side
#pragma unroll
        for (int r = 0; r < NRB; ++r) {
            double zc[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) zc[e] = clipd(tv[r][e], lo[r][e], hi[r][e]);
            if (r == f1_r) {
                project_polygon(P, tv[r][0] - shift[0], tv[r][1] - shift[1], zc[0], zc[1]);
                zc[0] += shift[0]; zc[1] += shift[1];
                zf[0] = zc[0]; zf[1] = zc[1];
            }
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (r < nrb) rOut0[e * KP + r * 8] = fma(P.rho, 2.0 * zc[e] - tv[r][e], -q[r][e]);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) qmag[e] = max(qmag[e], __shfl_xor_sync(0xffffffffu, qmag[e], o));
            if (g == 0) S.Res[1][rg][lc0 + e] = qmag[e];
        }
        group_sync<kGroupThreads>(bar_id);
        uint32_t thr[2], thr_accept[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            uint32_t mm = S.Res[1][0][lc0 + e];
#pragma unroll
            for (int q4 = 1; q4 < kRowGroups; ++q4) mm = max(mm, S.Res[1][q4][lc0 + e]);
            const double scale = fmax(1.0, __hiloint2double((int)mm, 0));
            thr[e] = mag_hi(P.eps * scale * P.stop_factor);
            thr_accept[e] = P.accept_eps > 0 ? mag_hi(P.accept_eps * scale * P.stop_factor) : 0u;
        }

        bool done[2] = {!live[0], !live[1]};
        bool stalled = false;              // the pair stopped without converging (see below)
        bool accepted = false;             // ... but within the reference solver's own tolerance: applied, status 2
        uint32_t last[2] = {0xffffffffu, 0xffffffffu};   // residuals of the last iteration (high words)
        int window = 0;
        uint32_t checkpoint = 0xffffffffu; // residual (high word) one stall window ago
        int iters[2] = {0, 0};
        for (int k = 0; k < P.max_iter; ++k) {
            double acc[NRB][2];
#pragma unroll
            for (int r = 0; r < NRB; ++r) acc[r][0] = acc[r][1] = 0.0;
            ContractDispatch<NRB, NRB, 0, false>::run(nrb, acc, aBase, aStride, bBase, KB);

            // a scenario is one unit: its two axes share the polygon, so they stop together
            const bool pair_done = done[0] && done[1];
            uint32_t rd[2] = {0u, 0u};
            if (!pair_done) {
#pragma unroll
                for (int r = 0; r < NRB; ++r) {
                    if (r >= nrb) continue;
                    double zo[2], tn[2], zn[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) zo[e] = clipd(tv[r][e], lo[r][e], hi[r][e]);
                    if (r == f1_r) { zo[0] = zf[0]; zo[1] = zf[1]; }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double d = acc[r][e] - zo[e];
                        tn[e] = fma(P.alpha, d, tv[r][e]);
                        zn[e] = clipd(tn[e], lo[r][e], hi[r][e]);
                        rd[e] = max(rd[e], mag_hi(d));
                    }
                    if (r == f1_r) {
                        project_polygon(P, tn[0] - shift[0], tn[1] - shift[1], zn[0], zn[1]);
                        zn[0] += shift[0]; zn[1] += shift[1];
                        zf[0] = zn[0]; zf[1] = zn[1];
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        tv[r][e] = tn[e];
                        acc[r][e] = fma(two_rho, zn[e], fma(-P.rho, tn[e], -q[r][e]));
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) rd[e] = max(rd[e], __shfl_xor_sync(0xffffffffu, rd[e], o));
                if (g == 0) S.Res[0][rg][lc0 + e] = rd[e];
            }
            group_sync<kGroupThreads>(bar_id);

            bool all_mine = true;
            if (!pair_done) {
                uint32_t cd[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#pragma unroll
                    for (int r = 0; r < NRB; ++r)
                        if (r < nrb) rOut0[e * KP + r * 8] = acc[r][e];
                    cd[e] = S.Res[0][0][lc0 + e];
#pragma unroll
                    for (int q4 = 1; q4 < kRowGroups; ++q4) cd[e] = max(cd[e], S.Res[0][q4][lc0 + e]);
                }
                bool both = (cd[0] <= thr[0]) && (cd[1] <= thr[1]);
                last[0] = cd[0]; last[1] = cd[1];
                // A QP whose residual has lost less than 10 % over a stall window (32 iterations) has stopped making
                // progress at this penalty (the final-stance ticks, where the state has left the region the
                // formulation can steer back: every row is active and the residual creeps): stop the pair a few dozen
                // iterations later instead of max_iter later.  A healthy solve contracts by 0.8-0.9 per iteration,
                // i.e. by 30x or more over such a window.
                if (!both && ++window >= P.stall_window) {
                    window = 0;
                    const uint32_t cur = max(cd[0], cd[1]);
                    if (checkpoint != 0xffffffffu &&
                        __hiloint2double((int)cur, 0) > 0.9 * __hiloint2double((int)checkpoint, 0)) { stalled = true; both = true; }
                    checkpoint = cur;
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (live[e]) { ++iters[e]; done[e] = both; }
            }
            all_mine = done[0] && done[1];
            const bool stop = group_sync_and<kGroupThreads>(bar_id, all_mine);
            if (stop) break;
        }

        // outputs: projected solution, iterations, status
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = col0 + lc0 + e;
            if (!live[e]) continue;
#pragma unroll
            for (int r = 0; r < NRB; ++r) {
                if (r >= nrb) continue;
                double z = clipd(tv[r][e], lo[r][e], hi[r][e]);
                if (r == f1_r) z = zf[e];
                P.sol[(size_t)c * NP + row0 + r * 8] = z;
                if (P.t_out) P.t_out[(size_t)c * NP + row0 + r * 8] = tv[r][e];
            }
            if (rg == 0 && g == 0) {
                // 0: converged to eps.  Not converged (stalled or out of iterations): 2 when the last residuals of BOTH
                // axes pass the reference solver's own stopping tolerance (OSQP at cvxpy's eps 1e-5 returns such a
                // point as "solved" and the reference applies it, zmp_controller.py:792-805), else 1 (fallback).
                const bool tight = done[e] && !stalled;
                accepted = !tight && last[0] <= thr_accept[0] && last[1] <= thr_accept[1];
                if (P.iters) P.iters[c] = iters[e];
                if (P.status) P.status[c] = tight ? 0 : (accepted ? 2 : 1);
            }
        }
    }
}

// next states and first footstep of every scenario: J0 = (e_0 + fc - Pzx[0] x) / Pzu[0,0]  (:806-812)
// A scenario whose QP did not converge falls back like the reference does when OSQP returns nothing (:796-804):
// zero jerk, footstep := current swing-foot interpolation.
__global__ void k_herdt_apply(int B, int NP, int N, int m, double T, double hg, const double* __restrict__ sol,
                              const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ fc,
                              const double* __restrict__ air, const int32_t* __restrict__ status,
                              double* __restrict__ next_x, double* __restrict__ next_y, double* __restrict__ f1) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * B) return;
    const int b = idx >> 1, axis = idx & 1;
    const bool failed = status && (status[2 * b] == 1 || status[2 * b + 1] == 1);     // 2 = applied at the reference tolerance
    const double* s = (axis ? y : x) + (size_t)b * 3;
    const double x0 = s[0], x1 = s[1], x2 = s[2];
    const double c3 = T * T * T / 6.0, c2 = T * T / 2.0;
    const double z0 = sol[(size_t)idx * NP] + fc[(size_t)b * 2 + axis];
    const double pred = fma(c2 - hg, x2, fma(T, x1, x0));
    const double jerk = failed ? 0.0 : (z0 - pred) / (c3 - T * hg);
    double* out = (axis ? next_y : next_x) + (size_t)b * 3;
    out[0] = fma(c3, jerk, fma(c2, x2, fma(T, x1, x0)));
    out[1] = fma(c2, jerk, fma(T, x2, x1));
    out[2] = fma(T, jerk, x2);
    if (f1) f1[(size_t)b * 2 + axis] = m > 0 ? ((failed && air) ? air[(size_t)b * 2 + axis] : sol[(size_t)idx * NP + N]) : nan("");
}

// Between two ticks of a walk (zmp_controller.py:495-525): swing-foot interpolation towards the first free
// footstep, footstep commit at the end of a single support, the push, the histories and the running
// status / iteration totals.  One thread per scenario; explicitly rounded operations (no contraction) so that
// the result is the same as the element-wise host formulas.
__global__ void k_herdt_book(int B, int n, int sample, int has_step, double inv_steps, int commit, int push,
                             const double* __restrict__ impulse, const double* __restrict__ f1,
                             const int32_t* __restrict__ tick_status, const int32_t* __restrict__ tick_iters,
                             double* __restrict__ next_x, double* __restrict__ next_y, double* __restrict__ fc,
                             double* __restrict__ air, double* __restrict__ x_hist, double* __restrict__ y_hist,
                             double* __restrict__ foot_hist, int32_t* __restrict__ status, long long* __restrict__ iters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double fcb[2] = {fc[2 * b], fc[2 * b + 1]}, airb[2] = {air[2 * b], air[2 * b + 1]};
    if (has_step) {
#pragma unroll
        for (int a = 0; a < 2; ++a) airb[a] = __dadd_rn(airb[a], __dmul_rn(inv_steps, __dsub_rn(f1[2 * b + a], airb[a])));
    }
    if (commit) {
#pragma unroll
        for (int a = 0; a < 2; ++a) { fcb[a] = has_step ? f1[2 * b + a] : airb[a]; airb[a] = fcb[a]; }
    }
    if (push) next_y[3 * b + 1] = __dsub_rn(next_y[3 * b + 1], impulse[b]);
#pragma unroll
    for (int a = 0; a < 2; ++a) { fc[2 * b + a] = fcb[a]; air[2 * b + a] = airb[a]; }
    const size_t at = (size_t)b * n + sample;
#pragma unroll
    for (int k = 0; k < 3; ++k) { x_hist[at * 3 + k] = next_x[3 * b + k]; y_hist[at * 3 + k] = next_y[3 * b + k]; }
    foot_hist[at * 2] = fcb[0]; foot_hist[at * 2 + 1] = fcb[1];
    if (status) {
        // low 16 bits: ticks that fell back; high 16 bits: ticks applied at the reference solver's tolerance
        const int worst = (tick_status[2 * b] == 1 || tick_status[2 * b + 1] == 1) ? 1 : max(tick_status[2 * b], tick_status[2 * b + 1]);
        status[b] += worst == 1 ? 1 : (worst == 2 ? (1 << 16) : 0);
    }
    if (iters) iters[b] += max(tick_iters[2 * b], tick_iters[2 * b + 1]);
}

__global__ void k_herdt_start(int B, int n, const double* __restrict__ x0, const double* __restrict__ y0, const double* __restrict__ fc0,
                              double* __restrict__ x, double* __restrict__ y, double* __restrict__ fc, double* __restrict__ air,
                              double* __restrict__ x_hist, double* __restrict__ y_hist, double* __restrict__ foot_hist,
                              int32_t* __restrict__ status, long long* __restrict__ iters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t at = (size_t)b * n;
    for (int k = 0; k < 3; ++k) {
        x[3 * b + k] = x0[3 * b + k]; y[3 * b + k] = y0[3 * b + k];
        x_hist[at * 3 + k] = x0[3 * b + k]; y_hist[at * 3 + k] = y0[3 * b + k];
    }
    for (int a = 0; a < 2; ++a) {
        fc[2 * b + a] = fc0[2 * b + a]; air[2 * b + a] = fc0[2 * b + a];
        foot_hist[at * 2 + a] = fc0[2 * b + a];
    }
    if (status) status[b] = 0;
    if (iters) iters[b] = 0;
}

template <int NRB, int NG>
int launch_qp(const QpParams& P, int sm_count, cudaStream_t stream) {
    const size_t nfrag = (size_t)(P.NP / 8) * (P.NP / 4) * 32;
    const size_t smem = 16 + (size_t)NG * ((sizeof(QpGroup) + 15) & ~(size_t)15) +
                        sizeof(double) * ((size_t)NG * kGroupCols * column_pitch(P.NP) + nfrag);
    auto kernel = k_qp_tick<NRB, NG>;
    LIPM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned ctas = (P.num_units + NG - 1) / NG;
    const unsigned grid = ctas < (unsigned)sm_count ? ctas : (unsigned)sm_count;
    kernel<<<grid, 32 * kRowGroups * NG, smem, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

}  // namespace
}  // namespace lipm

// ------------------------------------------------------------------------------------------------ C ABI
using namespace lipm;

struct herdt_plan {
    herdt_plan_desc_t desc;
    int N, sm_count;
    double *d_pzu, *d_pvu, *d_t, *d_v, *d_W;
};

struct herdt_timeline {
    herdt_plan* plan;
    int n_ticks, m_pad, n, NP;
    int8_t *d_phase, *d_kind;
    int32_t* h_m;           // host copy of m per tick
    double* d_frag;         // [n_ticks, nfrag]
    size_t nfrag;
    // per-tick scratch, grown on demand
    int cap_cols;
    double *d_q, *d_lo, *d_hi, *d_t0, *d_sol, *d_tprev;
    int prev_tick, prev_cols;   // what d_tprev holds (-1: nothing)
    unsigned int* d_counter;    // work-unit scheduler of k_qp_tick (per timeline: timelines may run concurrently)
};

extern "C" {

int herdt_plan_create(herdt_plan_t** out, const herdt_plan_desc_t* desc) {
    if (!out || !desc) { set_error("herdt_plan_create: null argument"); return LIPM_E_INVALID; }
    *out = nullptr;
    herdt_plan_desc_t d = *desc;
    if (d.horizon < 2 || d.horizon > kMaxResidentHorizon - 8) {
        set_error("herdt: horizon %d out of range [2, %d] (matrix + footstep slots must fit one SM)", d.horizon, kMaxResidentHorizon - 8);
        return LIPM_E_UNSUPPORTED;
    }
    if (!(d.h > 0) || !(d.g > 0) || !(d.dt > 0) || !(d.alpha > 0) || !(d.beta >= 0) || !(d.gamma >= 0)) {
        set_error("herdt: h, g, dt, alpha must be positive and beta, gamma non-negative");
        return LIPM_E_INVALID;
    }
    if (!(d.rho > 0)) d.rho = 6.0;     // tools/herdt_tune.py: (6, 1.8) needs 62 iterations per joint solve, (8, 1.6) 75
    if (!(d.relax > 0)) d.relax = 1.8;
    if (!(d.eps > 0)) d.eps = 1e-12;
    if (d.max_iter <= 0) d.max_iter = 600;
    if (d.stall_window <= 0) d.stall_window = 32;
    if (d.accept_eps == 0) d.accept_eps = 1e-5;       // cvxpy's eps_abs = eps_rel for OSQP
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { set_error("no CUDA device available (this library has no CPU path)"); return LIPM_E_CUDA; }
    if (d.device < 0 || d.device >= count) { set_error("device %d out of range", d.device); return LIPM_E_INVALID; }
    LIPM_CUDA_TRY(cudaSetDevice(d.device));
    cudaDeviceProp prop;
    LIPM_CUDA_TRY(cudaGetDeviceProperties(&prop, d.device));
    if (prop.major < 10) { set_error("device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor); return LIPM_E_UNSUPPORTED; }
    herdt_plan* p = new (std::nothrow) herdt_plan();
    if (!p) { set_error("out of host memory"); return LIPM_E_INVALID; }
    std::memset(p, 0, sizeof(*p));
    p->desc = d;
    p->N = d.horizon;
    p->sm_count = prop.multiProcessorCount;
    const int N = p->N;
    LIPM_CUDA_TRY(cudaMalloc(&p->d_pzu, sizeof(double) * N));
    LIPM_CUDA_TRY(cudaMalloc(&p->d_pvu, sizeof(double) * N));
    LIPM_CUDA_TRY(cudaMalloc(&p->d_t, sizeof(double) * N));
    LIPM_CUDA_TRY(cudaMalloc(&p->d_v, sizeof(double) * N));
    LIPM_CUDA_TRY(cudaMalloc(&p->d_W, sizeof(double) * N * N));
    k_herdt_toeplitz<<<(N + 127) / 128, 128>>>(N, d.dt, d.h / d.g, p->d_pzu, p->d_pvu);
    k_herdt_tv<<<1, 256, sizeof(double) * 2 * N>>>(N, p->d_pzu, p->d_pvu, p->d_t, p->d_v);
    k_herdt_w<<<dim3((N + 127) / 128, N), 128>>>(N, d.alpha, d.beta, p->d_t, p->d_v, p->d_W);
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(cudaDeviceSynchronize());
    *out = p;
    return LIPM_OK;
}

void herdt_plan_destroy(herdt_plan_t* p) {
    if (!p) return;
    void* ptrs[] = {p->d_pzu, p->d_pvu, p->d_t, p->d_v, p->d_W};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
}

int herdt_plan_export_w(const herdt_plan_t* p, double* host_out, size_t capacity) {
    if (!p || !host_out || capacity < (size_t)p->N * p->N) { set_error("herdt_plan_export_w: bad argument"); return LIPM_E_INVALID; }
    LIPM_CUDA_TRY(cudaMemcpy(host_out, p->d_W, sizeof(double) * p->N * p->N, cudaMemcpyDeviceToHost));
    return LIPM_OK;
}

int herdt_timeline_create(herdt_plan_t* plan, int n_ticks, int m_max, const int8_t* phase_of_row, const int8_t* row_kind,
                          const int32_t* m_per_tick, herdt_timeline_t** out) {
    if (!plan || !out || !phase_of_row || !row_kind || !m_per_tick || n_ticks < 1 || m_max < 0) {
        set_error("herdt_timeline_create: bad argument");
        return LIPM_E_INVALID;
    }
    *out = nullptr;
    const int N = plan->N;
    const int m_pad = m_max < 1 ? 1 : m_max;
    const int n = N + m_pad, NP = padded8(n);
    if (NP > kMaxResidentHorizon || m_pad > 15) {
        set_error("herdt: %d jerks + %d footsteps do not fit the resident-matrix kernel (limit %d)", N, m_pad, kMaxResidentHorizon);
        return LIPM_E_UNSUPPORTED;
    }
    herdt_timeline* t = new (std::nothrow) herdt_timeline();
    if (!t) { set_error("out of host memory"); return LIPM_E_INVALID; }
    std::memset(t, 0, sizeof(*t));
    t->plan = plan; t->n_ticks = n_ticks; t->m_pad = m_pad; t->n = n; t->NP = NP;
    t->prev_tick = -1;
    t->nfrag = (size_t)(NP / 8) * (NP / 4) * 32;
    t->h_m = new int32_t[n_ticks];
    std::memcpy(t->h_m, m_per_tick, sizeof(int32_t) * n_ticks);
    const size_t rows = (size_t)n_ticks * N;
    LIPM_CUDA_TRY(cudaMalloc(&t->d_phase, rows));
    LIPM_CUDA_TRY(cudaMalloc(&t->d_kind, rows));
    LIPM_CUDA_TRY(cudaMemcpy(t->d_phase, phase_of_row, rows, cudaMemcpyHostToDevice));
    LIPM_CUDA_TRY(cudaMemcpy(t->d_kind, row_kind, rows, cudaMemcpyHostToDevice));
    LIPM_CUDA_TRY(cudaMalloc(&t->d_frag, sizeof(double) * t->nfrag * n_ticks));
    double *d_P = nullptr, *d_inv = nullptr;
    int* d_info = nullptr;
    const size_t nn = (size_t)n * n;
    LIPM_CUDA_TRY(cudaMalloc(&d_P, sizeof(double) * nn * n_ticks));
    LIPM_CUDA_TRY(cudaMalloc(&d_inv, sizeof(double) * nn * n_ticks));
    LIPM_CUDA_TRY(cudaMalloc(&d_info, sizeof(int)));
    LIPM_CUDA_TRY(cudaMemset(d_info, 0, sizeof(int)));
    k_herdt_assemble<<<n_ticks, 256, sizeof(double) * N * m_pad>>>(N, m_pad, plan->desc.gamma, plan->d_W, t->d_phase, t->d_kind, d_P);
    k_shift_diag_batched<<<dim3((n + 127) / 128, n_ticks), 128>>>(n, plan->desc.rho, d_P);
    k_cholesky_batched<<<n_ticks, 1024>>>(n, d_P, d_info);
    k_cholesky_inverse_batched<<<dim3((n + 63) / 64, n_ticks), 64>>>(n, d_P, d_inv);
    k_pack_fragments_batched<<<dim3((unsigned)((t->nfrag + 255) / 256), n_ticks), 256>>>(n, NP, d_inv, t->d_frag);
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(cudaDeviceSynchronize());
    int info = 0;
    LIPM_CUDA_TRY(cudaMemcpy(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_P); cudaFree(d_inv); cudaFree(d_info);
    if (info != 0) {
        set_error("herdt: Cholesky of a tick Hessian broke down at pivot %d", info - 1);
        herdt_timeline_destroy(t);
        return LIPM_E_NUMERIC;
    }
    *out = t;
    return LIPM_OK;
}

void herdt_timeline_destroy(herdt_timeline_t* t) {
    if (!t) return;
    void* ptrs[] = {t->d_phase, t->d_kind, t->d_frag, t->d_q, t->d_lo, t->d_hi, t->d_t0, t->d_sol, t->d_tprev, t->d_counter};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete[] t->h_m;
    delete t;
}

int herdt_tick(herdt_timeline_t* t, int tick, int B, const double* x, const double* y, const double* fc,
               const double* air, const double* spread, const double* vref, int vref_batch, int n_ext, int foot_side,
               const double* poly_verts, int n_verts, int warm, double* next_x, double* next_y, double* f1,
               double* solution, int32_t* status, int32_t* iters, void* stream_) {
    if (!t || tick < 0 || tick >= t->n_ticks) { set_error("herdt_tick: bad timeline / tick"); return LIPM_E_INVALID; }
    if (B < 0 || !x || !y || !fc || !vref || !next_x || !next_y) { set_error("herdt_tick: missing argument"); return LIPM_E_INVALID; }
    if (B == 0) return LIPM_OK;
    if (vref_batch != 1 && vref_batch != B) { set_error("herdt_tick: vref_batch must be 1 or B"); return LIPM_E_INVALID; }
    if (foot_side != 0 && foot_side != 1) { set_error("herdt_tick: foot_side must be 0 (left) or 1 (right)"); return LIPM_E_INVALID; }
    const int m = t->h_m[tick];
    if (m > 0 && (!poly_verts || n_verts < 3 || n_verts > kMaxPolyVerts)) { set_error("herdt_tick: polygon with 3..%d vertices required", kMaxPolyVerts); return LIPM_E_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    herdt_plan* plan = t->plan;
    const herdt_plan_desc_t& d = plan->desc;
    const int N = plan->N, NP = t->NP, ncols = 2 * B;
    if (ncols > t->cap_cols) {
        // grown in stream order: cudaMalloc / cudaFree synchronise the whole device, which serialised the walks of
        // other timelines running on their own streams (9 timeline groups: 1.1 s became 2.5 - 3.3 s)
        void* ptrs[] = {t->d_q, t->d_lo, t->d_hi, t->d_t0, t->d_sol, t->d_tprev};
        for (void* q : ptrs) if (q) cudaFreeAsync(q, stream);
        t->d_q = t->d_lo = t->d_hi = t->d_t0 = t->d_sol = t->d_tprev = nullptr;
        t->cap_cols = 0;
        const size_t bytes = sizeof(double) * (size_t)ncols * NP;
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_q), bytes, stream));
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_lo), bytes, stream));
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_hi), bytes, stream));
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_t0), bytes, stream));
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_sol), bytes, stream));
        LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_tprev), bytes, stream));
        t->cap_cols = ncols;
        t->prev_tick = -1;
    }
    // warm start only from the tick just before, for the same batch
    const bool use_prev = warm != 0 && tick > 0 && t->prev_tick == tick - 1 && t->prev_cols == ncols;
    TickParams P;
    P.N = N; P.m_pad = t->m_pad; P.NP = NP; P.m = m; P.B = B; P.tick = tick; P.n_ext = n_ext; P.vref_batch = vref_batch;
    P.foot_side = foot_side;
    P.T = d.dt; P.hg = d.h / d.g; P.alpha = d.alpha; P.beta = d.beta;
    P.half[0] = 0.5 * d.foot_length; P.half[1] = 0.5 * d.foot_width; P.foot_spread = d.foot_spread;
    P.spread = spread;
    P.phase = t->d_phase + (size_t)tick * N; P.kind = t->d_kind + (size_t)tick * N;
    P.tcol = plan->d_t; P.vcol = plan->d_v;
    P.x = x; P.y = y; P.fc = fc; P.vref = vref;
    P.q = t->d_q; P.lo = t->d_lo; P.hi = t->d_hi; P.t0 = t->d_t0;
    P.t_prev = use_prev ? t->d_tprev : nullptr;
    P.shift_f = warm == 2 ? 1 : 0;
    P.m_prev = use_prev ? t->h_m[tick - 1] : 0;
    k_herdt_prepare<<<ncols, 128, sizeof(double) * 6 * N, stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());

    QpParams Q;
    Q.ncols = ncols; Q.NP = NP; Q.n = t->n; Q.f1_row = m > 0 ? N : -1;
    Q.max_iter = d.max_iter; Q.rho = d.rho; Q.alpha = d.relax; Q.eps = d.eps;
    Q.stop_factor = 1.0 / std::fmax(1.0 + d.relax, d.rho * d.relax);
    Q.accept_eps = d.accept_eps; Q.stall_window = d.stall_window;
    Q.mfrag = t->d_frag + (size_t)tick * t->nfrag;
    Q.q = t->d_q; Q.lo = t->d_lo; Q.hi = t->d_hi; Q.t0 = t->d_t0; Q.sol = solution ? solution : t->d_sol;
    Q.status = status; Q.iters = iters;
    Q.t_out = t->d_tprev;        // prepare has consumed the previous iterate (stream order)
    t->prev_tick = tick; t->prev_cols = ncols;
    Q.poly_shift = fc;
    Q.n_verts = m > 0 ? n_verts : 0;
    for (int i = 0; i < kMaxPolyVerts; ++i) {
        Q.verts[i][0] = (m > 0 && i < n_verts) ? poly_verts[2 * i] : 0.0;
        Q.verts[i][1] = (m > 0 && i < n_verts) ? poly_verts[2 * i + 1] : 0.0;
    }
    if (!t->d_counter) LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&t->d_counter), sizeof(unsigned int), stream));
    Q.tile_counter = t->d_counter;
    Q.num_units = (unsigned)((ncols + kGroupCols - 1) / kGroupCols);
    LIPM_CUDA_TRY(cudaMemsetAsync(t->d_counter, 0, sizeof(unsigned int), stream));
    int rc;
    switch ((NP / 8 + kRowGroups - 1) / kRowGroups) {
        case 1: rc = launch_qp<1, 4>(Q, plan->sm_count, stream); break;
        case 2: rc = launch_qp<2, 4>(Q, plan->sm_count, stream); break;
        case 3: rc = launch_qp<3, 4>(Q, plan->sm_count, stream); break;
        case 4: rc = launch_qp<4, 3>(Q, plan->sm_count, stream); break;
        case 5: rc = NP <= 152 ? launch_qp<5, 3>(Q, plan->sm_count, stream)      // 185 KB matrix + 3 groups
                               : launch_qp<5, 2>(Q, plan->sm_count, stream); break;   // NP = 160: 205 KB matrix, 2 groups
        default: set_error("herdt: no kernel variant for padded size %d", NP); return LIPM_E_UNSUPPORTED;
    }
    if (rc != LIPM_OK) return rc;
    k_herdt_apply<<<(ncols + 127) / 128, 128, 0, stream>>>(B, NP, N, m, d.dt, d.h / d.g, Q.sol, x, y, fc, air, status,
                                                           next_x, next_y, f1);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}

int herdt_rollout(herdt_timeline_t* t, int B, int n, const double* x0, const double* y0, const double* fc0,
                  const double* spread, const double* vref, int vref_batch, int n_ext, const int8_t* foot_side,
                  const int32_t* steps_left, const uint8_t* commit, const uint8_t* new_run, const double* poly_left,
                  int n_left, const double* poly_right, int n_right, int warm_start, const double* impulse, int force_tick,
                  double* x_hist, double* y_hist, double* foot_hist, int32_t* status, long long* iters, void* stream_) {
    if (!t || B < 0 || n < 1 || n - 1 > t->n_ticks) { set_error("herdt_rollout: bad timeline / sizes"); return LIPM_E_INVALID; }
    if (!x0 || !y0 || !fc0 || !vref || !foot_side || !steps_left || !commit || !new_run || !x_hist || !y_hist || !foot_hist) {
        set_error("herdt_rollout: missing argument");
        return LIPM_E_INVALID;
    }
    if (B == 0) return LIPM_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    // walk state of the batch: x, y (two copies, alternating), fc, air, f1, per-tick status / iterations
    const size_t doubles = (size_t)B * (4 * 3 + 3 * 2);
    unsigned char* scratch = nullptr;
    LIPM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), doubles * sizeof(double) + (size_t)B * 4 * sizeof(int32_t), stream));
    double* xs[2] = {reinterpret_cast<double*>(scratch), reinterpret_cast<double*>(scratch) + (size_t)B * 3};
    double* ys[2] = {xs[1] + (size_t)B * 3, xs[1] + (size_t)B * 6};
    double* fc = ys[1] + (size_t)B * 3;
    double* air = fc + (size_t)B * 2;
    double* f1 = air + (size_t)B * 2;
    int32_t* tick_status = reinterpret_cast<int32_t*>(f1 + (size_t)B * 2);
    int32_t* tick_iters = tick_status + (size_t)B * 2;
    const unsigned blocks = (unsigned)((B + 127) / 128);
    k_herdt_start<<<blocks, 128, 0, stream>>>(B, n, x0, y0, fc0, xs[0], ys[0], fc, air, x_hist, y_hist, foot_hist, status, iters);
    int rc = LIPM_OK;
    for (int i = 0; i < n - 1 && rc == LIPM_OK; ++i) {
        const int cur = i & 1, nxt = cur ^ 1;
        const int side = foot_side[i];
        const int has_step = t->h_m[i] > 0;
        const int warm = (i == 0 || !warm_start) ? 0 : (new_run[i] ? 2 : 1);
        rc = herdt_tick(t, i, B, xs[cur], ys[cur], fc, air, spread, vref, vref_batch, n_ext, side,
                        has_step ? (side ? poly_right : poly_left) : nullptr, has_step ? (side ? n_right : n_left) : 0, warm,
                        xs[nxt], ys[nxt], f1, nullptr, tick_status, tick_iters, stream_);
        if (rc != LIPM_OK) break;
        k_herdt_book<<<blocks, 128, 0, stream>>>(B, n, i + 1, has_step, has_step ? 1.0 / (double)steps_left[i] : 0.0, commit[i] ? 1 : 0,
                                                 (impulse != nullptr && i == force_tick) ? 1 : 0, impulse, f1, tick_status, tick_iters,
                                                 xs[nxt], ys[nxt], fc, air, x_hist, y_hist, foot_hist, status, iters);
    }
    if (rc == LIPM_OK && cudaGetLastError() != cudaSuccess) { set_error("herdt_rollout: launch failed"); rc = LIPM_E_CUDA; }
    cudaFreeAsync(scratch, stream);
    return rc;
}

int herdt_timeline_size(const herdt_timeline_t* t) { return t ? t->NP : 0; }

}  // extern "C"
