// Processing order of the scenarios of a batch.
//
// A solver group iterates its 8 columns until the slowest one has converged, so columns that
// need similar iteration counts should share a group.  What drives the count is known before
// the walk starts: the walk length and the push (when, how hard).  Scenarios are therefore
// processed in the order of the key (n_steps, force_tick / q, impulse): push ticks are binned so that a
// group's columns are close in BOTH push time and push size (with q = 1 a group of 8 out of the ~19
// scenarios that share a tick spans 40 % of the impulse range; measured on the 4096-scenario push sweep:
// q = 1: 19.7 M solves/s, q = 8..16: 21.4 M).  The balance point moves with the density of the batch,
// q ~ sqrt(group width x tick span / B), so it is computed on the device from the batch itself.  Results are written to
// the scenario's own slots, so the order is invisible to the caller and -- every column being
// solved independently and deterministically -- does not change a single bit of the output.
#include "lipm_internal.cuh"

#include <climits>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

namespace lipm {

namespace {

// span[0] = smallest, span[1] = largest push tick of the batch (span[0] > span[1]: no push at all)
__global__ void k_tick_span(int B, const int32_t* __restrict__ force_tick, int* __restrict__ span) {
    __shared__ int lo_s[32], hi_s[32];
    int lo = INT_MAX, hi = INT_MIN;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int t = force_tick[b];
        if (t >= 0) { lo = min(lo, t); hi = max(hi, t); }
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { lo_s[threadIdx.x >> 5] = lo; hi_s[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = min(lo, lo_s[w]); hi = max(hi, hi_s[w]); }
        span[0] = lo; span[1] = hi;
    }
}

__global__ void k_order_keys(int B, int n_uniform, const int32_t* __restrict__ n_steps,
                             const double* __restrict__ impulse, const int32_t* __restrict__ force_tick, int tick_bin_fixed,
                             int width, const int* __restrict__ span,
                             unsigned long long* __restrict__ keys, int* __restrict__ ids) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int tick_bin = tick_bin_fixed;
    if (tick_bin <= 0) {
        // q = 16.6 sqrt(width * span / B): 11 for 8-column groups on the 4096-scenario, 220-tick push sweep
        const int ticks = span ? span[1] - span[0] + 1 : 1;
        const float q = 16.6f * sqrtf((float)width * (float)(ticks > 1 ? ticks : 1) / (float)B);
        tick_bin = q < 1.f ? 1 : (q > 64.f ? 64 : (int)(q + 0.5f));
    }
    const unsigned long long n = (unsigned long long)min(max(n_steps ? n_steps[b] : n_uniform, 0), (1 << 20) - 1);
    long long tick = (impulse && force_tick) ? (long long)force_tick[b] + 1 : 0;   // 0 = no push
    tick = tick < 0 ? 0 : (tick > (1 << 20) - 1 ? (1 << 20) - 1 : tick);
    if (tick > 0) tick = 1 + (tick - 1) / tick_bin;
    const float f = impulse ? fabsf((float)impulse[b]) : 0.0f;
    const unsigned long long mag = (unsigned long long)(__float_as_uint(f) >> 8);   // 24 bits, order preserving
    keys[b] = (n << 44) | ((unsigned long long)tick << 24) | mag;
    ids[b] = b;
}

// Expected work of a unit of `width` consecutive columns (column c -> axis c / B, slot c % B): the group
// iterates every tick until its slowest column has converged, and what makes a column slow is its push --
// harder and earlier pushes keep more box rows active for longer.  Proxy: walk length + the largest
// impulse * remaining-ticks product of the unit (fits the measured iteration counts with R^2 = 0.89).
__global__ void k_unit_work(int num_units, int width, int ncols, int B, const int* __restrict__ order, int n_uniform,
                            const int32_t* __restrict__ n_steps, const double* __restrict__ impulse,
                            const int32_t* __restrict__ force_tick, float* __restrict__ keys, int* __restrict__ ids) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= num_units) return;
    float work = 0.f;
    for (int k = 0; k < width; ++k) {
        const int c = u * width + k;
        if (c >= ncols) break;
        const int axis = c / B;
        int b = c - axis * B;
        if (order) b = order[b];
        const int n = n_steps ? n_steps[b] : n_uniform;
        float w = (float)n;
        if (axis == 1 && impulse && force_tick && force_tick[b] >= 0 && force_tick[b] < n)
            w += 18.f * fabsf((float)impulse[b]) * (float)(n - force_tick[b]);
        work = fmaxf(work, w);
    }
    keys[u] = -work;          // ascending sort of the negated work = longest unit first
    ids[u] = u;
}

}  // namespace

// Dispatch order of the units (longest expected first, so that the tail of the launch is filled with short
// units): unit_order[i] = unit handed out i-th.  The result lives in `out` (released by its destructor).
int build_unit_order(int num_units, int width, int ncols, int B, const int* order, int n_uniform, const int32_t* n_steps,
                     const double* impulse, const int32_t* force_tick, StreamBuffer& out, cudaStream_t stream) {
    size_t temp_bytes = 0;
    LIPM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (float*)nullptr, (float*)nullptr, (int*)nullptr, (int*)nullptr,
                                                  num_units, 0, 32, stream));
    StreamBuffer keys, ids_in, temp;
    LIPM_CUDA_TRY(keys.alloc(sizeof(float) * 2 * (size_t)num_units, stream));
    LIPM_CUDA_TRY(ids_in.alloc(sizeof(int) * (size_t)num_units, stream));
    LIPM_CUDA_TRY(out.alloc(sizeof(int) * (size_t)num_units, stream));
    LIPM_CUDA_TRY(temp.alloc(temp_bytes, stream));
    float* keys_in = keys.as<float>();
    k_unit_work<<<(num_units + 127) / 128, 128, 0, stream>>>(num_units, width, ncols, B, order, n_uniform, n_steps, impulse,
                                                             force_tick, keys_in, ids_in.as<int>());
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp.ptr, temp_bytes, keys_in, keys_in + num_units, ids_in.as<int>(), out.as<int>(),
                                                  num_units, 0, 32, stream));
    return LIPM_OK;
}

// perm[slot] = scenario processed in that slot, in `out` (stream-ordered memory, released by its destructor).
int build_processing_order(int B, int n_uniform, const int32_t* n_steps, const double* impulse,
                           const int32_t* force_tick, int width, StreamBuffer& out, cudaStream_t stream) {
    size_t temp_bytes = 0;
    LIPM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                  (int*)nullptr, (int*)nullptr, B, 0, 64, stream));
    StreamBuffer keys, ids_in, temp, span;
    LIPM_CUDA_TRY(keys.alloc(sizeof(unsigned long long) * 2 * (size_t)B, stream));
    LIPM_CUDA_TRY(ids_in.alloc(sizeof(int) * (size_t)B, stream));
    LIPM_CUDA_TRY(out.alloc(sizeof(int) * (size_t)B, stream));
    LIPM_CUDA_TRY(temp.alloc(temp_bytes, stream));
    unsigned long long* keys_in = keys.as<unsigned long long>();
    int tick_bin = 0;                                     // 0: from the batch (k_order_keys)
#ifdef LIPM_DEBUG_KNOBS
    if (const char* env = std::getenv("LIPM_ORDER_TICK_BIN")) {      // measurement knob (tools/admm_tune.py)
        const int v = std::atoi(env);
        if (v >= 1 && v <= 4096) tick_bin = v;
    }
#endif
    if (impulse && force_tick && tick_bin <= 0) {
        LIPM_CUDA_TRY(span.alloc(2 * sizeof(int), stream));
        k_tick_span<<<1, 1024, 0, stream>>>(B, force_tick, span.as<int>());
    }
    k_order_keys<<<(B + 255) / 256, 256, 0, stream>>>(B, n_uniform, n_steps, impulse, force_tick,
                                                      span.ptr ? 0 : (tick_bin > 0 ? tick_bin : 1), width, span.as<int>(), keys_in,
                                                      ids_in.as<int>());
    LIPM_CUDA_TRY(cudaGetLastError());
    LIPM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp.ptr, temp_bytes, keys_in, keys_in + B, ids_in.as<int>(), out.as<int>(), B, 0, 64,
                                                  stream));
    return LIPM_OK;
}

}  // namespace lipm
