// Device side of libraytracegr_cuda: the scene constants, the warp-level scheduler pieces and the
// kernel bodies.  Included by raytracegr_cuda.cu (built-in metrics, compiled by nvcc) and, verbatim,
// by the run-time compiled program of a user-supplied metric (rtgr_metric_compile, NVRTC), so that
// both run the SAME integrator, event and colouring code.
#pragma once
#include "rtgr_trace.cuh"

namespace rtgr_dev {

using rtgr::Counters;
using rtgr::Job;
using rtgr::SceneConst;

__constant__ SceneConst c_scene;
__constant__ rtgr::StageTab c_tab = rtgr::make_stage_tab();

#ifndef RTGR_BLOCK_THREADS
#define RTGR_BLOCK_THREADS 128
#endif
#ifndef RTGR_MIN_BLOCKS
#define RTGR_MIN_BLOCKS 4   /* 128 registers/thread -> 4 warps per scheduler */
#endif
#ifndef RTGR_FETCH_CHUNK
#define RTGR_FETCH_CHUNK 32   /* >= 32: one refill never needs more than one new chunk */
#endif
constexpr int BLOCK_THREADS = RTGR_BLOCK_THREADS;
constexpr int MIN_BLOCKS_PER_SM = RTGR_MIN_BLOCKS;

// Stage accelerations of one thread: a column of shared memory, 7 stages x 2 x double2, laid out
// [stage][half][thread] so that a warp's 16-byte accesses are contiguous (conflict-free).
struct SmemAcc {
    using Backing = SmemAcc;
    double2* base;   // &smem[threadIdx.x]
    __device__ __forceinline__ SmemAcc backing() const { return *this; }
    __device__ __forceinline__ void load(int i, double v[4]) const {
        const double2 a = base[(2 * i) * BLOCK_THREADS], b = base[(2 * i + 1) * BLOCK_THREADS];
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    __device__ __forceinline__ void store(int i, const double v[4]) {
        base[(2 * i) * BLOCK_THREADS] = make_double2(v[0], v[1]);
        base[(2 * i + 1) * BLOCK_THREADS] = make_double2(v[2], v[3]);
    }
};

// ---------------------------------------------------------------------------------------------
// warp-level scheduler pieces used by rtgr::trace_loop
// ---------------------------------------------------------------------------------------------
struct WarpSched {
    unsigned long long* next;
    long long total;                         // ordinals in the queue (for the drain diagnostic only)
    int shared = 0;                          // the head lives in another GPU's memory or is drawn from by other
                                             // GPUs (cross-GPU tile queue, rtgr_frame): system-scope atomics
    unsigned long long t_empty = ~0ull;      // globaltimer when this warp first drew past the end
    // The warp draws ordinals from the global queue in private chunks of RTGR_FETCH_CHUNK (one 8x4-pixel
    // patch by default) and hands them to its lanes as they fall idle: the lanes of a warp then always
    // work on the same or on consecutive patches, whose rays take nearly the same number of steps, so
    // that they tend to finish (and start) in the same pass and share the out-of-line start-up and
    // finalisation code.  Also one global atomic per 32 rays instead of one per refill.
    long long c_base = 0;
    int c_left = 0;
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p); }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
    // Every lane calls this; lanes with want == true receive distinct queue ordinals.
    __device__ __forceinline__ int64_t fetch(bool want) {
        const unsigned m = __ballot_sync(0xffffffffu, want);
        if (m == 0) return -1;
        const int lane = threadIdx.x & 31;
        const int n = __popc(m);
        const int rank = __popc(m & ((1u << lane) - 1u));
        const int old = c_left < n ? c_left : n;          // served from what is left of the current chunk
        const long long old_base = c_base;
        c_base += old; c_left -= old;
        long long ord = old_base + rank;
        if (old < n) {                                    // warp-uniform: draw the next chunk
            unsigned long long nb = 0;
            if (lane == 0)
                nb = shared ? atomicAdd_system(next, (unsigned long long)RTGR_FETCH_CHUNK)
                            : atomicAdd(next, (unsigned long long)RTGR_FETCH_CHUNK);
            nb = __shfl_sync(0xffffffffu, nb, 0);
            if (t_empty == ~0ull && (long long)(nb + RTGR_FETCH_CHUNK) > total)
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_empty));
            if (rank >= old) ord = (long long)nb + (rank - old);
            c_base = (long long)nb + (n - old);
            c_left = RTGR_FETCH_CHUNK - (n - old);
        }
        return want ? int64_t(ord) : int64_t(-1);
    }
};

template <int METRIC, int RFORM, bool PATHS = false>
__device__ __forceinline__ void trace_kernel_body(const Job& job, unsigned long long* next, unsigned long long* counters) {
    __shared__ double2 s_acc[14 * BLOCK_THREADS];   // 28 KB per block
    WarpSched sched{next, job.total, job.queue_scope};
    SmemAcc acc{s_acc + threadIdx.x};
    Counters cnt{0, 0, 0, 0};
    rtgr::trace_loop<METRIC, RFORM, WarpSched, SmemAcc, PATHS>(c_scene, c_tab, job, sched, acc, cnt);
    // per-warp reduction of the work counters, one atomic per counter per warp
    unsigned long long v[4] = {cnt.rays, cnt.attempts, cnt.accepted, cnt.rejected};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(counters + k, v[k]);
        // drain diagnostics: when did the first warp find the queue empty, when did the last warp end
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(counters + 5, now);
        atomicMin(counters + 4, sched.t_empty);
#ifdef RTGR_PASS_STATS
        atomicAdd(counters + 6, (unsigned long long)cnt.passes);
        atomicAdd(counters + 7, ((unsigned long long)cnt.init_passes << 32) + cnt.fin_passes);
#endif
    }
}

template <int METRIC, int RFORM>
__device__ __forceinline__ void rhs_kernel_body(const double* __restrict__ states, int64_t n, double* __restrict__ derivs) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double y[8], A[4];
#pragma unroll
    for (int c = 0; c < 8; ++c) y[c] = states[8 * i + c];
    rtgr::accel<METRIC, RFORM>(c_scene, y, A);
#pragma unroll
    for (int c = 0; c < 4; ++c) { derivs[8 * i + c] = y[4 + c]; derivs[8 * i + 4 + c] = A[c]; }
}

template <int METRIC, int RFORM>
__device__ __forceinline__ void canvas_kernel_body(double* __restrict__ pixels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= c_scene.ni) return;
    double x[4], u[4];
    rtgr::canvas_pixel<METRIC, RFORM>(c_scene, i, j, x, u);
    double* px = pixels + 11 * (int64_t(i) + int64_t(j) * c_scene.ni);
#pragma unroll
    for (int c = 0; c < 4; ++c) { px[c] = x[c]; px[4 + c] = u[c]; }
    px[8] = px[9] = px[10] = 0.0;
}

}  // namespace rtgr_dev
