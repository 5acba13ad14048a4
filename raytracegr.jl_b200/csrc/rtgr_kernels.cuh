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
// RGB8 patch staging (north-star item 3: "writes vectorised, coalesced RGB tiles").  A chunk of the queue is one
// 8x4-pixel patch = four 24-byte row segments of the image.  Its rays end in different passes, so the colours are
// collected in shared memory (96 bytes per patch, two patches in flight per warp) and the complete patch leaves
// in ONE store instruction: twelve lanes write 8 bytes each.  For the image that lives in another GPU's memory
// (rtgr_frame) that is 12 peer writes of 8 bytes per patch instead of 96 of one byte.
struct alignas(16) PatchStage {       // (16-byte multiples: the slots of consecutive warps stay aligned for the uint2 reads)
    uint32_t px[2][24];     // 32 pixels x 3 bytes, patch-lane order (l = 8*row + column)
    int32_t key[2];         // pixel index of the patch's first pixel, -1: slot free
    uint32_t mask[2];       // patch lanes whose colour has arrived
    int32_t last;           // the slot opened most recently
};
static_assert(sizeof(PatchStage) % 16 == 0, "PatchStage slots must keep 8-byte alignment in an array");

// Queue ordinal of a chunk (= one 8x4-pixel patch in render mode) -> pixel coordinates of the patch's first pixel
__device__ __forceinline__ void patch_origin(const Job& job, long long nb, int& pi0, int& pj0) {
    const int64_t m = nb >> 10;
    const int sub = int(nb & 1023) >> 5;
    int64_t t = job.tile_offset + m * job.tile_stride;
    if (job.tile_order) t = job.tile_order[t];
    const int ty = int(t / job.tiles_x), tx = int(t % job.tiles_x);
    pi0 = tx * RTGR_TILE_W + (sub & 3) * 8;
    pj0 = ty * RTGR_TILE_H + (sub >> 2) * 4;
}

// The Pixel structs of two chunks per warp (CHUNK_RAYS below): 32 x 11 doubles each, slot l = ordinal - chunk base.
// In canvas mode (rgb written in place into the Pixel array, Job::stage_canvas) a slot is also where the colours
// of the chunk's rays are collected, and the complete patch goes back as four whole rows of 704 contiguous bytes.
// 5664 bytes per warp: with the 28 KB of stage accelerations more than the 48 KB a block may have statically, so
// this is DYNAMIC shared memory (the host passes CHUNK_SMEM_BYTES and raises the kernel's limit).
struct alignas(16) ChunkSlots {
    double px[2][11 * RTGR_FETCH_CHUNK];
    int32_t key[2];         // canvas mode: pixel index of the patch's first pixel while its colours are collected, else -1
    uint32_t mask[2];       // patch lanes whose colour has arrived
    int32_t last;           // the slot of the chunk drawn most recently
    int32_t pad[3];
};
static_assert(sizeof(ChunkSlots) % 16 == 0, "ChunkSlots must keep 16-byte alignment in an array");
constexpr int CHUNK_SMEM_BYTES = int(sizeof(ChunkSlots)) * (BLOCK_THREADS / 32);

// STAGE_RGB8: the RGB8 image is written through the patch staging above.
// CHUNK_RAYS: the rays come from a Pixel array (Job::pixels_in -- device memory, or the caller's page-locked HOST
// canvas read and written over PCIe).  The warp then reads the 32 rays of a chunk TOGETHER when it draws the chunk:
// 2816 bytes (the patch's four rows of 8 pixels x 88 bytes, or 32 consecutive pixels of a 1-D array) in fully coalesced
// 256-byte instructions into shared memory, instead of 64 scattered bytes per ray at the moment a lane needs one; and
// in canvas mode the colours go back the same way, as whole rows of the patch (pos and normal rewritten with the
// values read), instead of 24 bytes per ray into the middle of an 88-byte struct.  Measured with 8 GPUs on one host
// canvas (profiles/r02w_*): the partial-line writes of 8 x 30 M rays/s are what the host memory system cannot absorb
// (trace kernel 38.3 ms against 35.0 ms with the image in GPU memory; 36.3 ms patch-wise); a flat 8K canvas at one
// GPU: 208 -> 174 ms from the reads alone, 162 ms with the write-back.
template <bool STAGE_RGB8, bool CHUNK_RAYS = false>
struct WarpSchedT {
    static constexpr bool STAGE = STAGE_RGB8;
    static constexpr bool PREFETCH = CHUNK_RAYS;
    __device__ static __forceinline__ ChunkSlots* cslots() {       // (dynamic shared memory: see ChunkSlots)
        extern __shared__ __align__(16) unsigned char rtgr_dyn_smem[];
        return reinterpret_cast<ChunkSlots*>(rtgr_dyn_smem) + (threadIdx.x >> 5);
    }
    __device__ static __forceinline__ void init_cslots() {
        ChunkSlots* cs = cslots();
        if ((threadIdx.x & 31) == 0) { cs->key[0] = cs->key[1] = -1; cs->mask[0] = cs->mask[1] = 0u; cs->last = 0; }
        __syncwarp();
    }
    unsigned long long* next;
    long long total;                         // ordinals in the queue (for the drain diagnostic only)
    int shared = 0;                          // the head lives in another GPU's memory or is drawn from by other
                                             // GPUs (cross-GPU tile queue, rtgr_frame): system-scope atomics
    // STAGE: this warp's staging slots.  Found again from the thread index wherever they are needed (rare code) rather
    // than kept in the scheduler object: a pointer member would be two more registers live through the step loop,
    // which has none to spare.
    __device__ static __forceinline__ PatchStage* slots() {
        __shared__ PatchStage s_stage[BLOCK_THREADS / 32];
        return &s_stage[threadIdx.x >> 5];
    }
    __device__ static __forceinline__ void init_slots() {
        PatchStage* st = slots();
        if ((threadIdx.x & 31) == 0) { st->key[0] = st->key[1] = -1; st->mask[0] = st->mask[1] = 0u; st->last = 0; }
        __syncwarp();
    }
    unsigned long long t_empty = ~0ull;      // globaltimer when this warp first drew past the end
    // The warp draws ordinals from the global queue in private chunks of RTGR_FETCH_CHUNK (one 8x4-pixel
    // patch by default) and hands them to its lanes as they fall idle: the lanes of a warp then always
    // work on the same or on consecutive patches, whose rays take nearly the same number of steps, so
    // that they tend to finish (and start) in the same pass and share the out-of-line start-up and
    // finalisation code.  Also one global atomic per 32 rays instead of one per refill.
    long long c_base = 0;
    int c_left = 0;
    long long c_new = -1;                    // STAGE / PREFETCH: the chunk the last fetch drew (live only within the refill block)
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p); }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
    // Every lane calls this; lanes with want == true receive distinct queue ordinals.
    __device__ __forceinline__ int64_t fetch(bool want, const Job& job) {
        const unsigned m = __ballot_sync(0xffffffffu, want);
        if (STAGE || PREFETCH) c_new = -1;
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
            if (STAGE || PREFETCH) c_new = (long long)nb < total ? (long long)nb : -1;
        }
        return want ? int64_t(ord) : int64_t(-1);
    }

    // ---- RGB8 patch staging.  All of it is rare code kept OUT OF LINE: the step loop is as large as the
    // ---- instruction cache lets it be (inlined, these 6 KB cost the 4K frame 2 %).
    // A new chunk = a new patch: give it a staging slot if it lies wholly inside the image (border patches are
    // stored directly).  With both slots still collecting, the older one is written out as far as it got and its
    // remaining rays store directly (rare: one ray of a patch outliving a whole later patch).  Warp-uniform.
    __device__ static __forceinline__ void open_patch(long long nb, const Job& job) {
        PatchStage* st = slots();
        int pi0, pj0;
        patch_origin(job, nb, pi0, pj0);
        if (pi0 + 8 > c_scene.ni || pj0 + 4 > c_scene.nj) return;
        __syncwarp();
        int s = (st->key[0] < 0) ? 0 : ((st->key[1] < 0) ? 1 : -1);
        if (s < 0) {
            s = 1 - st->last;
            const int l = threadIdx.x & 31;
            if ((st->mask[s] >> l) & 1u) {
                const uint8_t* b = reinterpret_cast<const uint8_t*>(st->px[s]) + 3 * l;
                const int64_t pix = int64_t(st->key[s]) + (l & 7) + int64_t(l >> 3) * c_scene.ni;
                rtgr::store_rgb8_direct(job, pix, uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16));
            }
        }
        __syncwarp();      // every lane has read the slot table before lane 0 rewrites it
        if ((threadIdx.x & 31) == 0) { st->key[s] = pi0 + pj0 * c_scene.ni; st->mask[s] = 0u; st->last = s; }
        __syncwarp();
    }
    // The staging work of a refill, in ONE out-of-line call (every call site inside the step loop costs it code):
    // write out the slots that are complete -- twelve lanes, 8 bytes each (row l/3 of the patch, piece l%3 of its 24
    // bytes) -- and then open a slot for the chunk that has just been drawn, if any.
    __device__ static __noinline__ void stage_work(const Job& job, bool s0, bool s1, long long nb) {
        PatchStage* st = slots();
        __syncwarp();
        const int l = threadIdx.x & 31;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!(s == 0 ? s0 : s1)) continue;
            if (l < 12) {
                const int row = l / 3, part = l - 3 * row;
                const uint2 v = *reinterpret_cast<const uint2*>(&st->px[s][6 * row + 2 * part]);
                uint8_t* o = job.rgb8 + 3 * (int64_t(st->key[s]) + int64_t(row) * c_scene.ni) + 8 * part;
                *reinterpret_cast<uint2*>(o) = v;
            }
            __syncwarp();
            if (l == 0) { st->key[s] = -1; st->mask[s] = 0u; }
        }
        __syncwarp();
        if (nb >= 0) open_patch(nb, job);
    }
    // Right after the fetch of a refill block: `code` is -2 / -3 on a lane whose ray has just completed slot 0 / 1.
    __device__ __forceinline__ void stage_refill(const Job& job, int code) {
        const bool s0 = any(code == -2), s1 = any(code == -3);
        if (s0 || s1 || c_new >= 0) stage_work(job, s0, s1, c_new);
    }
    // ---- PREFETCH: the rays of a chunk read together, the colours of a canvas patch written together ----
    // Write slot s back: the patch's four rows, 88 doubles each, three coalesced 256-byte instructions per row.
    __device__ static __forceinline__ void write_back(const Job& job, ChunkSlots* cs, int s) {
        const int lane = threadIdx.x & 31;
        double* canvas = job.rgb_f64 - 8;          // the Pixel array itself (stage_canvas: rgb_f64 = pixels + 8, stride 11)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double* dst = canvas + 11 * (int64_t(cs->key[s]) + int64_t(r) * c_scene.ni);
#pragma unroll
            for (int k = 0; k < 3; ++k) { const int i = lane + 32 * k; if (i < 88) dst[i] = cs->px[s][88 * r + i]; }
        }
    }
    // The chunk work of a refill in ONE out-of-line call: write back the slots whose patches are complete, then give
    // the chunk that has just been drawn (`nb` >= 0) a slot -- a free one, else the older one, whose patch is written
    // back as far as it got (its remaining rays store their colours directly, later in program order) -- and read
    // its 32 Pixel structs into it.  Warp-uniform.
    __device__ static __noinline__ void chunk_work(const Job& job, bool s0, bool s1, long long nb) {
        ChunkSlots* cs = cslots();
        const int lane = threadIdx.x & 31;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!(s == 0 ? s0 : s1)) continue;
            write_back(job, cs, s);
            __syncwarp();
            if (lane == 0) { cs->key[s] = -1; cs->mask[s] = 0u; }
        }
        __syncwarp();
        if (nb < 0) return;
        const int s = 1 - cs->last;                  // the slots alternate: the new chunk takes the older one
        if (cs->key[s] >= 0) { write_back(job, cs, s); __syncwarp(); }
        double* buf = cs->px[s];
        int key = -1;
        if (job.mode == rtgr::JOB_PIXELS) {          // 32 consecutive pixels of the 1-D array
            const long long left = job.total - nb;
            const int nd = int(left < RTGR_FETCH_CHUNK ? left : RTGR_FETCH_CHUNK) * 11;
            const double* src = job.pixels_in + 11 * nb;
#pragma unroll
            for (int k = 0; k < 11; ++k) { const int i = lane + 32 * k; if (i < nd) buf[i] = src[i]; }
        } else {                                     // four rows of 8 pixels: 88 contiguous doubles each
            int pi0, pj0;
            patch_origin(job, nb, pi0, pj0);
            const int w = c_scene.ni - pi0, h = c_scene.nj - pj0;     // columns / rows of the patch inside the canvas
            const int nd = (w < 8 ? (w > 0 ? w : 0) : 8) * 11;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r >= h) break;
                const double* src = job.pixels_in + 11 * (int64_t(pi0) + int64_t(pj0 + r) * c_scene.ni);
#pragma unroll
                for (int k = 0; k < 3; ++k) { const int i = lane + 32 * k; if (i < nd) buf[88 * r + i] = src[i]; }
            }
            // colours are collected only for a patch that lies wholly inside the canvas
            if (job.stage_canvas && w >= 8 && h >= 4) key = pi0 + pj0 * c_scene.ni;
        }
        __syncwarp();      // every lane has read the slot table before lane 0 rewrites it
        if (lane == 0) { cs->key[s] = key; cs->mask[s] = 0u; cs->last = s; }
        __syncwarp();
    }
    // Right after the fetch of a refill block: hands lane `want` the ray of its ordinal (pos, normal -> v[0..7]);
    // `code` is -2 / -3 on a lane whose ray has just completed the patch of slot 0 / 1.  Lanes still served from the
    // previous chunk take their rays before anything else happens (that chunk's slot is `last`).
    __device__ __forceinline__ void take_rays(const Job& job, bool want, int64_t ord, int code, double v[8]) {
        ChunkSlots* cs = cslots();
        const long long cn = c_new;
        const bool valid = want && ord < total;
        const bool from_new = valid && cn >= 0 && ord >= cn;
        const int l11 = 11 * int(ord & (RTGR_FETCH_CHUNK - 1));
        if (valid && !from_new) {
            const double* mine = cs->px[cs->last] + l11;
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = mine[c];
        }
        const bool s0 = any(code == -2), s1 = any(code == -3);
        if (s0 || s1 || cn >= 0) {
            chunk_work(job, s0, s1, cn);
            if (from_new) {
                const double* mine = cs->px[cs->last] + l11;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = mine[c];
            }
        }
    }
    // One finished ray's colour in canvas mode (called by that lane alone, from divergent code inside finalize_ray):
    // into the slot that collects its patch, else straight into the canvas.  Returns the slot the ray completed, else -1.
    __device__ static __forceinline__ int put_rgbf(const SceneConst& sc, const Job& job, int32_t pix, const double col[3]) {
        ChunkSlots* cs = cslots();
        const int pj = pix / sc.ni, pi = pix - pj * sc.ni;
        const int key = (pi & ~7) + (pj & ~3) * sc.ni;
        const int s = (cs->key[0] == key) ? 0 : ((cs->key[1] == key) ? 1 : -1);
        if (s < 0) {
            for (int c = 0; c < 3; ++c) job.rgb_f64[int64_t(job.rgb_stride) * pix + c] = col[c];
            return -1;
        }
        const int l = (pi & 7) + ((pj & 3) << 3);
        double* o = cs->px[s] + 11 * l + 8;
        o[0] = col[0]; o[1] = col[1]; o[2] = col[2];      // (read only behind the __syncwarp of a later chunk_work: no fence)
        const unsigned bit = 1u << l;
        return ((atomicOr(&cs->mask[s], bit) | bit) == 0xffffffffu) ? s : -1;
    }

    // One finished ray's colour (called by that lane alone, from divergent code inside finalize_ray).  Returns the
    // slot the ray completed, else -1.  (No fence: the bytes are READ only behind the __syncwarp of a later stage_work,
    // which orders them; the mask is only ever touched atomically or behind that barrier.)
    __device__ static __forceinline__ int put_rgb8(const SceneConst& sc, const Job& job, int32_t pix, uint32_t rgb) {
        PatchStage* st = slots();
        const int pj = pix / sc.ni, pi = pix - pj * sc.ni;
        const int key = (pi & ~7) + (pj & ~3) * sc.ni;
        const int s = (st->key[0] == key) ? 0 : ((st->key[1] == key) ? 1 : -1);
        if (s < 0) { rtgr::store_rgb8_direct(job, pix, rgb); return -1; }
        const unsigned bit = 1u << ((pi & 7) + ((pj & 3) << 3));
        uint8_t* b = reinterpret_cast<uint8_t*>(st->px[s]) + 3 * ((pi & 7) + ((pj & 3) << 3));
        b[0] = uint8_t(rgb); b[1] = uint8_t(rgb >> 8); b[2] = uint8_t(rgb >> 16);
        return ((atomicOr(&st->mask[s], bit) | bit) == 0xffffffffu) ? s : -1;
    }
};

// STAGE: the launch writes a tile-ordered RGB8 image whose 24-byte row segments are 8-byte aligned (the host checks:
// stage_rgb8_wanted) through the patch staging above.
template <int METRIC, int RFORM, bool PATHS = false, bool STAGE = false, bool PREFETCH = false>
__device__ __forceinline__ void trace_kernel_body(const Job& job, unsigned long long* next, unsigned long long* counters) {
    __shared__ double2 s_acc[14 * BLOCK_THREADS];   // 28 KB per block
    WarpSchedT<STAGE, PREFETCH> sched{next, job.total, job.queue_scope};
    static_assert(!(STAGE || PREFETCH) || RTGR_FETCH_CHUNK == 32, "staging and chunk reads need chunk = patch = 32 rays");
    if (STAGE) WarpSchedT<STAGE, PREFETCH>::init_slots();
    if (PREFETCH) WarpSchedT<STAGE, PREFETCH>::init_cslots();
    SmemAcc acc{s_acc + threadIdx.x};
    Counters cnt{0, 0, 0, 0};
    rtgr::trace_loop<METRIC, RFORM, WarpSchedT<STAGE, PREFETCH>, SmemAcc, PATHS>(c_scene, c_tab, job, sched, acc, cnt);
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
