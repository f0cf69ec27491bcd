// Device-side footstep plans and ZMP bound series for parameter sweeps.
//
// Restates, one thread per walk variant, the reference's host generators:
//   generate_footsteps          generators/footstep_generator.py:19-49
//   CoPGenerator.generate_cop_trajectory   generators/cop_generator.py:73-115 (state machine :80-101)
// The gait timeline is driven by a floating-point clock that accumulates t += dt and compares it with an
// accumulated switching time; the irregular phase lengths this produces (24- and 25-sample single supports
// for configs/default.json) are part of what the controller sees.  Every operation below is the same IEEE
// double operation in the same order (explicitly rounded intrinsics: no FMA contraction), so the series are
// bit-identical to the host generators, which are pinned against the reference itself.
#include "lipm_internal.cuh"

namespace lipm {
namespace {

constexpr int kMaxFootsteps = 256;
constexpr int kStanding = 0, kDouble = 1, kSingle = 2;

struct GenParams {
    int V, n_max;
    const double *distance, *step_length, *foot_spread;
    double ssp, dsp, standing, dt;
    double half_len, half_wid;      // sole 0.11 x 0.05 (hard-coded by the reference, footstep_generator.py:34)
    double *z_min, *z_max;          // [V, n_max, 2]
    int8_t* states;                 // [V, n_max] or nullptr
    int32_t* n_out;                 // [V]: samples produced (negative: the series needs -n_out > n_max samples)
};

__global__ void k_generate_bounds(const GenParams P) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.V) return;
    const double distance = P.distance[v], step = P.step_length[v], spread = P.foot_spread[v];
    double fx[kMaxFootsteps], fy[kMaxFootsteps];
    int count = 0;
    fx[count] = 0.0; fy[count++] = -spread;
    fx[count] = 0.0; fy[count++] = spread;
    double x = 0.0, side = spread;
    while (x < distance && count < kMaxFootsteps - 1) {
        const double remaining = __dsub_rn(distance, x);
        const double half_step = __dmul_rn(0.5, step);
        const double stride = remaining <= step ? (remaining < half_step ? remaining : half_step) : step;   // min(remaining, 0.5 step)
        x = __dadd_rn(x, stride);
        side = -side;
        fx[count] = x; fy[count++] = side;
    }
    fx[count] = x; fy[count++] = -side;

    const int last = count - 1;
    int current = 1, state = kStanding, n = 0;
    double clock = 0.0, switch_at = P.standing;
    double* zmax = P.z_max + (size_t)v * P.n_max * 2;
    double* zmin = P.z_min + (size_t)v * P.n_max * 2;
    int8_t* st = P.states ? P.states + (size_t)v * P.n_max : nullptr;
    while (current < count) {
        if (clock > switch_at) {
            const bool on_last = current == last;
            const bool terminal = state == kStanding && on_last;
            double extra = 0.0;
            int consumed = 0;
            if (state == kStanding) {
                if (on_last) consumed = 1;
                else { state = kDouble; extra = P.dsp; }
            } else if (state == kSingle) {
                state = kDouble; extra = P.dsp; consumed = 1;
            } else {
                if (on_last) { state = kStanding; extra = P.standing; }
                else { state = kSingle; extra = P.ssp; }
            }
            if (!terminal) switch_at = __dadd_rn(switch_at, extra);
            current += consumed;
        }
        if (current < count) {
            const double lx_hi = __dadd_rn(fx[current], P.half_len), lx_lo = __dsub_rn(fx[current], P.half_len);
            const double ly_hi = __dadd_rn(fy[current], P.half_wid), ly_lo = __dsub_rn(fy[current], P.half_wid);
            double hx = lx_hi, hy = ly_hi, lx = lx_lo, ly = ly_lo;
            if (state != kSingle) {
                const double tx_hi = __dadd_rn(fx[current - 1], P.half_len), tx_lo = __dsub_rn(fx[current - 1], P.half_len);
                const double ty_hi = __dadd_rn(fy[current - 1], P.half_wid), ty_lo = __dsub_rn(fy[current - 1], P.half_wid);
                hx = fmax(tx_hi, lx_hi); hy = fmax(ty_hi, ly_hi);
                lx = fmin(tx_lo, lx_lo); ly = fmin(ty_lo, ly_lo);
            }
            if (n < P.n_max) {
                zmax[n * 2] = hx; zmax[n * 2 + 1] = hy;
                zmin[n * 2] = lx; zmin[n * 2 + 1] = ly;
                if (st) st[n] = (int8_t)state;
            }
            ++n;
        }
        clock = __dadd_rn(clock, P.dt);
    }
    P.n_out[v] = n <= P.n_max ? n : -n;
}

}  // namespace
}  // namespace lipm

using namespace lipm;

extern "C" int lipm_generate_bounds(int V, const double* distance, const double* step_length, const double* foot_spread,
                                    double ssp_duration, double dsp_duration, double standing_duration, double dt,
                                    int n_max, double* z_min, double* z_max, int8_t* states, int32_t* n_out, void* stream) {
    if (V < 0 || n_max < 1 || !distance || !step_length || !foot_spread || !z_min || !z_max || !n_out) {
        set_error("lipm_generate_bounds: bad argument");
        return LIPM_E_INVALID;
    }
    if (!(dt > 0) || !(ssp_duration >= 0) || !(dsp_duration >= 0) || !(standing_duration >= 0)) {
        set_error("lipm_generate_bounds: durations must be non-negative and dt positive");
        return LIPM_E_INVALID;
    }
    if (V == 0) return LIPM_OK;
    GenParams P;
    P.V = V; P.n_max = n_max;
    P.distance = distance; P.step_length = step_length; P.foot_spread = foot_spread;
    P.ssp = ssp_duration; P.dsp = dsp_duration; P.standing = standing_duration; P.dt = dt;
    P.half_len = 0.11 / 2; P.half_wid = 0.05 / 2;
    P.z_min = z_min; P.z_max = z_max; P.states = states; P.n_out = n_out;
    k_generate_bounds<<<(V + 63) / 64, 64, 0, (cudaStream_t)stream>>>(P);
    LIPM_CUDA_TRY(cudaGetLastError());
    return LIPM_OK;
}
