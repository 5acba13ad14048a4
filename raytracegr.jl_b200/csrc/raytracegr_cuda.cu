// libraytracegr_cuda: CUDA kernels (sm_100a) + the C ABI of include/raytracegr_cuda.h.
//
// Kernel design:
//   * persistent grid, one ray per thread; a lane that finishes pulls the next ray from a global
//     atomic queue (warp-aggregated: one atomicAdd per refill, __ballot_sync/__shfl_sync to hand
//     out the slots), so warps stay full through regions where step counts differ by 30x;
//   * the whole per-ray pipeline is fused: make_canvas (render mode) -> initial dt -> Tsit5
//     attempts with error control -> event detection on the dense output -> root-find ->
//     classification + colouring -> RGB store;  no per-ray state ever touches HBM;
//   * FP64 throughout, no tensor cores (the path is not a dense contraction).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rtgr_scene.h"
#include "rtgr_trace.cuh"
#include "rtgr_jit.h"

namespace {

using rtgr::Counters;
using rtgr::Job;
using rtgr::SceneConst;

static_assert(sizeof(rtgr_object) == 88 && sizeof(rtgr_params) == 72 && sizeof(rtgr_camera) == 136 &&
                  sizeof(rtgr_pixel) == 88 && sizeof(rtgr_stats) == 56, "ABI struct layout");

}  // namespace

#include "rtgr_kernels.cuh"

namespace {

using rtgr_dev::BLOCK_THREADS;
using rtgr_dev::MIN_BLOCKS_PER_SM;
using rtgr_dev::c_scene;

template <int METRIC, int RFORM>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
trace_kernel(Job job, unsigned long long* next, unsigned long long* counters) {
    rtgr_dev::trace_kernel_body<METRIC, RFORM>(job, next, counters);
}
template <int METRIC, int RFORM>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
trace_paths_kernel(Job job, unsigned long long* next, unsigned long long* counters) {
    rtgr_dev::trace_kernel_body<METRIC, RFORM, true>(job, next, counters);
}
// The same kernel with the RGB8 patch staging (rtgr_kernels.cuh, PatchStage): a warp's 8x4-pixel patch leaves in one
// 8-byte-per-lane store.  Launched when the image lives in ANOTHER GPU's memory (see stage_rgb8_wanted).
template <int METRIC, int RFORM>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
trace_stage_kernel(Job job, unsigned long long* next, unsigned long long* counters) {
    rtgr_dev::trace_kernel_body<METRIC, RFORM, false, true>(job, next, counters);
}
// ... and the kernel for rays that come from a Pixel array (rtgr_trace_pixels, rtgr_trace_canvas[_frame]): a warp reads
// the 32 rays of a chunk together, coalesced, when it draws the chunk (rtgr_kernels.cuh, WarpSchedT::load_chunk).
template <int METRIC, int RFORM>
__global__ void __launch_bounds__(BLOCK_THREADS, MIN_BLOCKS_PER_SM)
trace_pixels_kernel(Job job, unsigned long long* next, unsigned long long* counters) {
    rtgr_dev::trace_kernel_body<METRIC, RFORM, false, false, true>(job, next, counters);
}
template <int METRIC, int RFORM>
__global__ void rhs_kernel(const double* __restrict__ states, int64_t n, double* __restrict__ derivs) {
    rtgr_dev::rhs_kernel_body<METRIC, RFORM>(states, n, derivs);
}
template <int METRIC, int RFORM>
__global__ void canvas_kernel(double* __restrict__ pixels) {
    rtgr_dev::canvas_kernel_body<METRIC, RFORM>(pixels);
}

// Gather the pixels of a device's tiles out of its full-frame output buffer into a compact tile-major buffer
// (slot (m, q): tile m of the device's list, pixel q = 32*row + column of the tile): a device of a statically
// dealt multi-device call then returns only ITS tiles to the host, not the whole frame.  `elem` bytes per pixel.
__global__ void pack_tiles_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int32_t* __restrict__ list,
                                  int offset, int stride, long long count, int tiles_x, int ni, int nj, int elem) {
    for (long long m = blockIdx.x; m < count; m += gridDim.x) {
        const long long t = list ? (long long)list[m] : (long long)offset + m * stride;
        const int ty = int(t / tiles_x), tx = int(t % tiles_x);
        for (int q = threadIdx.x; q < RTGR_TILE_W * RTGR_TILE_H; q += blockDim.x) {
            const int i = tx * RTGR_TILE_W + (q % RTGR_TILE_W), j = ty * RTGR_TILE_H + (q / RTGR_TILE_W);
            uint8_t* b = dst + (size_t(m) * (RTGR_TILE_W * RTGR_TILE_H) + q) * elem;
            if (i >= ni || j >= nj) {       // the part of a border tile outside the image: defined bytes (never read back)
                for (int w = 0; w < elem; ++w) b[w] = 0;
                continue;
            }
            const uint8_t* a = src + (size_t(j) * ni + i) * elem;
            if ((elem & 3) == 0) {
                for (int w = 0; w < elem / 4; ++w) reinterpret_cast<uint32_t*>(b)[w] = reinterpret_cast<const uint32_t*>(a)[w];
            } else {
                for (int w = 0; w < elem; ++w) b[w] = a[w];
            }
        }
    }
}

// Register-resident DFMA chains: the FP64 roofline denominator.
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Operand-mix variants of the same register-resident chains (8 independent accumulators per thread).
// They measure how the FP64 pipe's issue rate depends on how many distinct 64-bit REGISTER operands
// an instruction reads (constants come from the constant bank / immediates and cost no register-file
// bandwidth).  Every mode reports "DFMA-equivalent" TFLOP/s = 2 x thread-instructions/s, so the
// numbers are directly comparable with mode 1 (the pipe limit).
//   1  DFMA a = a*C1 + C2        (1 register operand)          -- fp64_peak_kernel above
//   2  DFMA a = a*b + C          (2 register operands)
//   3  DFMA a = b*c + a          (3 distinct register operands, no operand shared between neighbours)
//   4  DMUL a = a*b              (2)
//   5  DADD a = a + b            (2)
//   6  DMUL a = a*C              (1)
//   7  DFMA a = b*b + a          (3 slots, 2 distinct registers)
//   8  DFMA a = b*c + a, neighbouring instructions share b in the same slot (operand-reuse friendly)
//   9  alternating mode-3 DFMA and mode-4 DMUL
template <int MODE>
__global__ void fp64_mix_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    // multipliers near 1 / addends near 0 keep every chain finite for any iteration count
    double b0 = 1.0 - 1e-9 * a0, b1 = 1.0 - 1e-9 * a1, b2 = 1.0 - 1e-9 * a2, b3 = 1.0 - 1e-9 * a3;
    double c0 = 1e-12 * a0, c1 = 1e-12 * a1, c2 = 1e-12 * a2, c3 = 1e-12 * a3;
    const double K = 1e-9, M1 = 0.999999;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 2) {
                a0 = fma(a0, b0, K); a1 = fma(a1, b1, K); a2 = fma(a2, b2, K); a3 = fma(a3, b3, K);
                a4 = fma(a4, b0, K); a5 = fma(a5, b1, K); a6 = fma(a6, b2, K); a7 = fma(a7, b3, K);
            } else if (MODE == 3) {
                a0 = fma(b0, c1, a0); a1 = fma(b1, c2, a1); a2 = fma(b2, c3, a2); a3 = fma(b3, c0, a3);
                a4 = fma(b0, c2, a4); a5 = fma(b1, c3, a5); a6 = fma(b2, c0, a6); a7 = fma(b3, c1, a7);
            } else if (MODE == 4) {
                a0 = __dmul_rn(a0, b0); a1 = __dmul_rn(a1, b1); a2 = __dmul_rn(a2, b2); a3 = __dmul_rn(a3, b3);
                a4 = __dmul_rn(a4, b0); a5 = __dmul_rn(a5, b1); a6 = __dmul_rn(a6, b2); a7 = __dmul_rn(a7, b3);
            } else if (MODE == 5) {
                a0 = __dadd_rn(a0, c0); a1 = __dadd_rn(a1, c1); a2 = __dadd_rn(a2, c2); a3 = __dadd_rn(a3, c3);
                a4 = __dadd_rn(a4, c0); a5 = __dadd_rn(a5, c1); a6 = __dadd_rn(a6, c2); a7 = __dadd_rn(a7, c3);
            } else if (MODE == 6) {
                a0 = __dmul_rn(a0, M1); a1 = __dmul_rn(a1, M1); a2 = __dmul_rn(a2, M1); a3 = __dmul_rn(a3, M1);
                a4 = __dmul_rn(a4, M1); a5 = __dmul_rn(a5, M1); a6 = __dmul_rn(a6, M1); a7 = __dmul_rn(a7, M1);
            } else if (MODE == 7) {
                a0 = fma(c0, c0, a0); a1 = fma(c1, c1, a1); a2 = fma(c2, c2, a2); a3 = fma(c3, c3, a3);
                a4 = fma(c0, c0, a4); a5 = fma(c1, c1, a5); a6 = fma(c2, c2, a6); a7 = fma(c3, c3, a7);
            } else if (MODE == 8) {
                a0 = fma(c0, b0, a0); a1 = fma(c0, b1, a1); a2 = fma(c0, b2, a2); a3 = fma(c0, b3, a3);
                a4 = fma(c1, b0, a4); a5 = fma(c1, b1, a5); a6 = fma(c1, b2, a6); a7 = fma(c1, b3, a7);
            } else {   // 9
                a0 = fma(b0, c1, a0); a1 = __dmul_rn(a1, b1); a2 = fma(b2, c3, a2); a3 = __dmul_rn(a3, b3);
                a4 = fma(b0, c2, a4); a5 = __dmul_rn(a5, b1); a6 = fma(b2, c0, a6); a7 = __dmul_rn(a7, b3);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Issue-mix probe: full-rate DFMAs (modes 10, 11: two register operands; mode 12: three) interleaved with
// independent 32-bit integer instructions (1 per DFMA in modes 10 and 12, 3 per DFMA in mode 11).  The
// integer work runs on another pipe; what the modes show is whether it takes issue / register-read
// bandwidth away from the FP64 instructions.  Reported as DFMA-equivalent TFLOP/s of the DFMAs alone.
template <int MODE>
__global__ void fp64_int_mix_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b0 = 1.0 - 1e-9 * a0, b1 = 1.0 - 1e-9 * a1, b2 = 1.0 - 1e-9 * a2, b3 = 1.0 - 1e-9 * a3;
    double c0 = 1e-12 * a0, c1 = 1e-12 * a1, c2 = 1e-12 * a2, c3 = 1e-12 * a3;
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    unsigned j0 = blockIdx.x | 1u, j1 = j0 + 2, j2 = j0 + 4, j3 = j0 + 6;
    const double K = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 12) {
                a0 = fma(b0, c1, a0); a1 = fma(b1, c2, a1); a2 = fma(b2, c3, a2); a3 = fma(b3, c0, a3);
                a4 = fma(b0, c2, a4); a5 = fma(b1, c3, a5); a6 = fma(b2, c0, a6); a7 = fma(b3, c1, a7);
            } else {
                a0 = fma(a0, b0, K); a1 = fma(a1, b1, K); a2 = fma(a2, b2, K); a3 = fma(a3, b3, K);
                a4 = fma(a4, b0, K); a5 = fma(a5, b1, K); a6 = fma(a6, b2, K); a7 = fma(a7, b3, K);
            }
            i0 = i0 * j0 + j1; i1 = i1 * j1 + j2; i2 = i2 * j2 + j3; i3 = i3 * j3 + j0;
            i4 = i4 * j0 + j2; i5 = i5 * j1 + j3; i6 = i6 * j2 + j0; i7 = i7 * j3 + j1;
            if (MODE == 11) {
                i0 ^= i4 >> 3; i1 ^= i5 >> 3; i2 ^= i6 >> 3; i3 ^= i7 >> 3;
                i4 ^= i1 >> 5; i5 ^= i2 >> 5; i6 ^= i3 >> 5; i7 ^= i0 >> 5;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + double(i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7);
}

// Latency / parallelism probe: CHAINS independent dependent-DFMA chains per thread (1 register operand each,
// so no operand-read limit); launched with one block per SM and a chosen number of warps per scheduler.
// With 1 chain and 1 warp per scheduler the rate is 1/latency.
template <int CHAINS>
__global__ void fp64_chain_kernel(double* out, int iters, double seed) {
    double a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = seed + threadIdx.x + c;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 64 / CHAINS; ++k) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
thread_local std::string g_err;

int fail(const std::string& msg) { g_err = msg; return -1; }

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Device {
    int id = 0;
    int sm_count = 0;
    unsigned char uuid[16] = {0};      // the GPU's UUID: "is that frame in MY memory?" across processes (rtgr_frame_open)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* d_next = nullptr;      // queue head
    unsigned long long* d_counters = nullptr;  // 4 counters
    DevBuf pixels, rgb8, rgbf, fstate, objid, status, nsteps, scratch, order, pack;
    std::vector<int32_t> order_host;   // what `order` holds (upload_tile_list)
    DevBuf h_stage;  // pinned host staging
    std::vector<cudaEvent_t> chunk_ev;  // completion markers of the pieces of a chunked D2H copy
    int64_t resident_n = 0;
    bool launched = false;  // a trace kernel was launched on this device during the current call
    int grid[5] = {0, 0, 0, 0, 0};  // persistent grid size per kernel variant (variant_of)
    bool chunk_smem_set[5] = {false, false, false, false, false};   // trace_pixels_kernel: dynamic shared memory limit raised
};

}  // namespace

// A run-time compiled user metric (rtgr_metric_compile): the loaded library and its three kernels.
struct UserMetric {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_trace = nullptr, k_trace_stage = nullptr, k_trace_paths = nullptr, k_rhs = nullptr, k_canvas = nullptr;
    double par[16] = {0};
    int blocks_per_sm = 0;
    bool alive = false;
};

struct rtgr_ctx {
    std::vector<Device> devs;
    std::vector<UserMetric> metrics;   // metric id = RTGR_USER_METRIC_BASE + index
    // the cost-sorted tile list of the last frame geometry (sorting 32 400 tiles of an 8K frame takes
    // longer on the host than a millisecond; the keys of a repeated camera / canvas are identical)
    std::vector<double> order_keys;
    std::vector<int32_t> order_sorted;
    std::vector<struct rtgr_frame*> frames;   // open shared frames of this context (closed by rtgr_destroy)
    int peer_state = 0;                       // devices 1.. can use device 0's memory (loads, stores, atomics): 0 unknown, 1 yes, -1 no
    cudaEvent_t ev_shared = nullptr;          // "device 0's queue head and inputs are ready" (shared-queue launches)
};

// A frame shared by several GPUs (rtgr_frame_create / rtgr_frame_open): ONE allocation in the owner GPU's
// memory holding the tile-queue heads and the RGB8 image.  Every participating GPU -- other devices of this
// context through peer access, other processes through a CUDA IPC mapping -- draws its rays from the same
// queue with system-scope atomics and stores its pixels straight into the owner's image over NVLink.
//   byte   0: queue head of even frames     byte 128: queue head of odd frames
//   byte  64: magic, version, ni, nj        byte 256: RGB8 image, nj x ni x 3 (PNG order)
struct rtgr_frame {
    rtgr_ctx* ctx = nullptr;
    int ni = 0, nj = 0;
    bool owner = false;
    uint8_t* base = nullptr;
    int home = -1;               // CUDA device ordinal (in this process) the memory lives on, -1 if unknown
    unsigned long long epoch = 0;   // frames rendered through this handle
    int participants = 0;           // rtgr_frame_set_participants: how many GPUs work on the frame (0: not told)
    unsigned char owner_uuid[16] = {0};   // UUID of the GPU whose memory holds the frame (from the frame's header)
};

namespace {

constexpr size_t FRAME_HEADER = 256, FRAME_INFO = 64, FRAME_HEAD_STRIDE = 128;
constexpr uint32_t FRAME_MAGIC = 0x52544746u;   // "RTGF"
static_assert(sizeof(cudaIpcMemHandle_t) == RTGR_IPC_HANDLE_BYTES, "IPC handle size");

// Make memory on CUDA device `home` usable (loads, stores AND atomics) from device `dev` of this process.
// `enable` = false only checks: the device that opened an IPC mapping already has its (lazily enabled)
// peer access, and enabling it by hand would create a context on `home` in this process for nothing.
int enable_peer(int dev, int home, bool enable) {
    if (home < 0 || dev == home) return 0;
    int can = 0, atomics = 0;
    CU(cudaDeviceCanAccessPeer(&can, dev, home));
    if (!can)
        return fail("device " + std::to_string(dev) + " has no peer access to device " + std::to_string(home) +
                    " (a shared frame needs NVLink/NVSwitch peer memory)");
    CU(cudaDeviceGetP2PAttribute(&atomics, cudaDevP2PAttrNativeAtomicSupported, dev, home));
    if (!atomics)
        return fail("device " + std::to_string(dev) + " cannot do native atomics on device " + std::to_string(home) +
                    "'s memory (the shared tile queue needs NVLink atomics)");
    if (!enable) return 0;
    CU(cudaSetDevice(dev));
    const cudaError_t e = cudaDeviceEnablePeerAccess(home, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
}

// Several devices in ONE context work through a call's tiles from ONE queue (head, inputs and output buffers in
// device 0's memory, which the others reach as peer memory over NVLink -- the mechanism of rtgr_render_frame),
// so they finish together without a cost model and the result is already assembled in one place.  Boxes without
// peer access / native peer atomics between the devices, and RTGR_MULTI_QUEUE=static, use per-device tile sets.
bool multi_queue_shared(rtgr_ctx* ctx) {
    const char* e = getenv("RTGR_MULTI_QUEUE");
    if (e && e[0] == 's' && e[1] == 't') return false;      // "static"
    if (ctx->peer_state == 0) {
        ctx->peer_state = 1;
        for (size_t k = 1; k < ctx->devs.size(); ++k)
            if (enable_peer(ctx->devs[k].id, ctx->devs[0].id, true)) { ctx->peer_state = -1; break; }
        if (ctx->peer_state > 0) {
            cudaSetDevice(ctx->devs[0].id);
            if (cudaEventCreateWithFlags(&ctx->ev_shared, cudaEventDisableTiming) != cudaSuccess) ctx->peer_state = -1;
        }
    }
    return ctx->peer_state > 0;
}

int ensure(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}
int ensure_pinned(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr; b.cap = 0;
    CU(cudaMallocHost(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

// 0: Minkowski; 1, 2: Kerr-Schild with the radius as written / textbook; 3, 4: the same for a == 0 (the reference's
// own scene, src:276), where the right-hand side is half the work (ks_accel_a0)
constexpr int N_VARIANTS = 5;
int variant_of(const rtgr_params* p) {
    if (p->metric == RTGR_MINKOWSKI) return 0;
    return (p->r_formula == RTGR_R_AS_WRITTEN ? 1 : 2) + (p->a == 0.0 ? 2 : 0);
}

// params->metric >= RTGR_USER_METRIC_BASE selects a user metric of this context
int user_metric_of(rtgr_ctx* ctx, const rtgr_params* p, UserMetric** um) {
    *um = nullptr;
    if (!ctx || !p || p->metric < RTGR_USER_METRIC_BASE) return 0;
    const size_t idx = size_t(p->metric - RTGR_USER_METRIC_BASE);
    if (idx >= ctx->metrics.size() || !ctx->metrics[idx].alive) return fail("unknown user metric id (rtgr_metric_compile)");
    *um = &ctx->metrics[idx];
    return 0;
}

template <class F> int with_variant(int v, F&& f) {
    switch (v) {
        case 0: return f(std::integral_constant<int, RTGR_MINKOWSKI>{}, std::integral_constant<int, RTGR_R_AS_WRITTEN>{});
        case 1: return f(std::integral_constant<int, RTGR_KERR_SCHILD>{}, std::integral_constant<int, RTGR_R_AS_WRITTEN>{});
        case 2: return f(std::integral_constant<int, RTGR_KERR_SCHILD>{}, std::integral_constant<int, RTGR_R_CORRECTED>{});
        case 3: return f(std::integral_constant<int, RTGR_KERR_SCHILD>{}, std::integral_constant<int, rtgr::RFORM_A0 + RTGR_R_AS_WRITTEN>{});
        default: return f(std::integral_constant<int, RTGR_KERR_SCHILD>{}, std::integral_constant<int, rtgr::RFORM_A0 + RTGR_R_CORRECTED>{});
    }
}

int persistent_grid(Device& d, int variant) {
    if (d.grid[variant]) return d.grid[variant];
    int per_sm = 0;
    with_variant(variant, [&](auto M, auto R) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel<decltype(M)::value, decltype(R)::value>,
                                                      BLOCK_THREADS, 0);
        return 0;
    });
    if (per_sm < 1) per_sm = 1;
    d.grid[variant] = per_sm * d.sm_count;
    return d.grid[variant];
}

// RGB8 patch staging (north-star item 3, "vectorised, coalesced RGB tiles"): can this job's image be written as
// whole 8x4-pixel patches -- tile-ordered RGB8 output whose 24-byte row segments are 8-byte aligned -- and should
// it?  For an image in the GPU's OWN memory the L2 merges the byte stores anyway, so the staging kernel is used
// where every store crosses NVLink: `remote` = the image lives in another GPU's memory (5x fewer store sectors and
// 2.5x fewer NVLink bytes, profiles/r03*_peer_store_probe.jsonl).  RTGR_RGB8_STAGING=1 / 0 forces it on / off wherever
// it is possible (measurements, tests).
bool stage_rgb8_wanted(const Job& job, int ni, bool remote) {
    if (!job.rgb8 || job.mode != rtgr::JOB_RENDER || job.paths || (ni & 7) != 0 || (reinterpret_cast<uintptr_t>(job.rgb8) & 7) != 0)
        return false;
    if (const char* e = getenv("RTGR_RGB8_STAGING")) return e[0] == '1';
    return remote;
}

// Launch the trace kernel for `job` on device d (scene constants already uploaded).
// `queue` != nullptr: draw from that (shared, already initialised) queue head instead of the device's own.
int launch_trace(Device& d, int variant, const Job& job, UserMetric* um = nullptr, unsigned long long* queue = nullptr) {
    // a lane keeps its pixel index in 32 bits (2^31 rays = 189 GB of Pixel structs: more than one GPU holds)
    if (job.total >= (int64_t(1) << 31)) return fail("more than 2^31 - 1 rays in one launch");
    if (!queue) {
        CU(cudaMemsetAsync(d.d_next, 0, sizeof(unsigned long long), d.stream));
        queue = d.d_next;
    }
    CU(cudaMemsetAsync(d.d_counters, 0, 8 * sizeof(unsigned long long), d.stream));
    CU(cudaMemsetAsync(d.d_counters + 4, 0xff, sizeof(unsigned long long), d.stream));   // min-slot starts at ~0
    int grid = 0;
    if (um) {
        if (!um->blocks_per_sm) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)um->k_trace, BLOCK_THREADS, 0) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                per_sm = 2;
            }
            um->blocks_per_sm = per_sm;
        }
        grid = um->blocks_per_sm * d.sm_count;
    } else {
        grid = persistent_grid(d, variant);
    }
    const int64_t resident_rays = int64_t(grid) * BLOCK_THREADS;    // rays the GPU holds at full occupancy
    const int64_t lanes_needed = (job.total + BLOCK_THREADS - 1) / BLOCK_THREADS;
    if (lanes_needed < grid) grid = int(std::max<int64_t>(1, lanes_needed));
    // A frame with fewer rays than the GPU has threads is bound by its longest ray times the latency of one
    // step attempt, and that latency grows with the number of warps that share a scheduler's FP64 pipe: such
    // a frame runs with ONE CTA per SM (one warp per scheduler), the warps working through their patches one
    // after the other (example2 at 200x200: 3.57 ms against 4.41 ms with everything resident at once; from
    // about twice that many rays on, full occupancy wins: profiles/r01zz_small_frames.log).  Not for
    // Minkowski, whose steps are not FP64-pipe-bound.  RTGR_CTAS_PER_SM=k overrides the cap (experiments).
    int cta_cap = (variant != 0 && job.total <= resident_rays) ? 1 : 0;
    if (const char* e = getenv("RTGR_CTAS_PER_SM")) cta_cap = atoi(e);
    if (cta_cap >= 1 && int64_t(cta_cap) * d.sm_count < grid) grid = cta_cap * d.sm_count;
    // rays from a Pixel array are read chunk-wise (trace_pixels_kernel); RTGR_CHUNK_RAYS=0: ray by ray (measurements)
    // Measured (profiles/r02u_*): a flat 8K canvas in host memory 208 -> 174 ms; the 4K Kerr-Schild canvas, whose rays
    // take 450 steps each, 282.6 -> 283.3 ms (the reads were hidden anyway, the extra code is not) -- so by default
    // only for the metric whose rays are short.  RTGR_CHUNK_RAYS=1 / 0 forces it on / off.
    // A canvas (rgb in place, stage_canvas) also gets its colours back a patch at a time.  That is what lets MANY GPUs
    // work on one host canvas -- eight GPUs' 24-byte writes into the middle of 88-byte structs are more than the host
    // memory system absorbs: trace kernel 38.3 ms against 35.0 ms with the image in GPU memory (profiles/r02w_*) -- but
    // at one GPU the extra code costs 1 % (285.3 against 282.1 ms).  So: for Minkowski, and for a host canvas shared
    // by four or more GPUs (stage_canvas == 2).  RTGR_CHUNK_RAYS=1 / 0 forces it on / off.
    bool chunk_rays = job.pixels_in != nullptr && !job.paths && (variant == 0 || job.stage_canvas == 2) && !um;
    if (const char* e = getenv("RTGR_CHUNK_RAYS")) chunk_rays = job.pixels_in != nullptr && !job.paths && !um && e[0] != '0';
    CU(cudaEventRecord(d.ev0, d.stream));
    if (um) {
        Job j = job;
        void* args[] = {&j, &queue, &d.d_counters};
        CU(cudaLaunchKernel((const void*)(job.paths ? um->k_trace_paths : (job.stage_rgb8 ? um->k_trace_stage : um->k_trace)),
                            dim3(grid), dim3(BLOCK_THREADS), args, 0, d.stream));
    } else if (job.paths) {
        with_variant(variant, [&](auto M, auto R) {
            trace_paths_kernel<decltype(M)::value, decltype(R)::value><<<grid, BLOCK_THREADS, 0, d.stream>>>(job, queue, d.d_counters);
            return 0;
        });
    } else if (chunk_rays) {
        const int rc = with_variant(variant, [&](auto M, auto R) {
            auto kernel = trace_pixels_kernel<decltype(M)::value, decltype(R)::value>;
            if (!d.chunk_smem_set[variant]) {     // static + dynamic shared memory exceed the 48 KB a kernel gets by default;
                // four blocks of 51 KB per SM need the largest shared-memory carveout
                if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rtgr_dev::CHUNK_SMEM_BYTES) != cudaSuccess) return -1;
                if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return -1;
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK_THREADS, rtgr_dev::CHUNK_SMEM_BYTES) != cudaSuccess) return -1;
                if (getenv("RTGR_VERBOSE")) fprintf(stderr, "rtgr: trace_pixels_kernel variant %d: %d blocks per SM\n", variant, per_sm);
                d.chunk_smem_set[variant] = true;
            }
            kernel<<<grid, BLOCK_THREADS, rtgr_dev::CHUNK_SMEM_BYTES, d.stream>>>(job, queue, d.d_counters);
            return 0;
        });
        if (rc) { CU(cudaGetLastError()); return fail("cudaFuncSetAttribute(trace_pixels_kernel) failed"); }
    } else if (job.stage_rgb8) {
        with_variant(variant, [&](auto M, auto R) {
            trace_stage_kernel<decltype(M)::value, decltype(R)::value><<<grid, BLOCK_THREADS, 0, d.stream>>>(job, queue, d.d_counters);
            return 0;
        });
    } else {
        with_variant(variant, [&](auto M, auto R) {
            trace_kernel<decltype(M)::value, decltype(R)::value><<<grid, BLOCK_THREADS, 0, d.stream>>>(job, queue, d.d_counters);
            return 0;
        });
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(d.ev1, d.stream));
    d.launched = true;
    return 0;
}

// `sc` goes into the __constant__ block of the kernels that will run: the built-in ones, or the user
// metric's own copy inside its run-time compiled library (the current device must be d.id).
int upload_scene(Device& d, SceneConst& sc, UserMetric* um = nullptr) {
    if (um) {
        std::memcpy(sc.user_par, um->par, sizeof(sc.user_par));
        void* dptr = nullptr; size_t bytes = 0;
        CU(cudaLibraryGetGlobal(&dptr, &bytes, um->lib, "_ZN8rtgr_dev7c_sceneE"));
        if (bytes != sizeof(SceneConst)) return fail("user metric library: scene block size mismatch");
        CU(cudaMemcpyAsync(dptr, &sc, sizeof(SceneConst), cudaMemcpyHostToDevice, d.stream));
        return 0;
    }
    CU(cudaMemcpyToSymbolAsync(c_scene, &sc, sizeof(SceneConst), 0, cudaMemcpyHostToDevice, d.stream));
    return 0;
}

// The tile list a device works through (ordinal -> tile id).  It is the same from call to call for a repeated
// camera / canvas, so the copy is skipped when the device already holds it.
int upload_tile_list(Device& d, const std::vector<int32_t>& list) {
    if (d.order_host == list && d.order.p) return 0;
    if (ensure(d.order, list.size() * sizeof(int32_t))) return -1;
    CU(cudaMemcpyAsync(d.order.p, list.data(), list.size() * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
    CU(cudaStreamSynchronize(d.stream));     // (pageable source: done before `list` or the cached copy can change)
    d.order_host = list;
    return 0;
}

int collect_stats(rtgr_ctx* ctx, rtgr_stats* stats, double total_ms) {
    rtgr_stats s{};
    for (auto& d : ctx->devs) {
        if (!d.launched) continue;
        d.launched = false;
        CU(cudaSetDevice(d.id));
        unsigned long long h[8];
        CU(cudaMemcpyAsync(h, d.d_counters, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        s.rays += h[0];
        s.rhs_evals += 6ull * h[1] + 2ull * h[0];
        s.steps_accepted += h[2];
        s.steps_rejected += h[3];
        s.kernel_ms = std::max(s.kernel_ms, double(ms));
#ifdef RTGR_PASS_STATS
        fprintf(stderr, "[pass stats] warp passes %llu, with start-up code %llu (%.1f %%), with finalisation %llu (%.1f %%), rays %llu\n",
                h[6], h[7] >> 32, 100.0 * double(h[7] >> 32) / double(h[6]), h[7] & 0xffffffffull,
                100.0 * double(h[7] & 0xffffffffull) / double(h[6]), h[0]);
#endif
        if (h[4] != ~0ull && h[5] > h[4]) s.drain_ms = std::max(s.drain_ms, double(h[5] - h[4]) * 1e-6);
    }
    s.total_ms = total_ms;
    if (stats) *stats = s;
    return 0;
}

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// copy a device's packed tiles (pack_tiles_kernel's layout: tile m of its list, 32 x 32 slots) into the user's image
void scatter_packed_tiles(const uint8_t* src, uint8_t* dst, int ni, int nj, size_t elem_bytes, int tiles_x,
                          int offset, int stride, int64_t count, const int32_t* list) {
    for (int64_t m = 0; m < count; ++m) {
        const int64_t t = list ? int64_t(list[m]) : int64_t(offset) + m * stride;
        const int ty = int(t / tiles_x), tx = int(t % tiles_x);
        const int i0 = tx * RTGR_TILE_W, j0 = ty * RTGR_TILE_H;
        const int w = std::min(RTGR_TILE_W, ni - i0), hgt = std::min(RTGR_TILE_H, nj - j0);
        for (int j = 0; j < hgt; ++j)
            std::memcpy(dst + (size_t(j0 + j) * ni + i0) * elem_bytes,
                        src + (size_t(m) * (RTGR_TILE_W * RTGR_TILE_H) + size_t(j) * RTGR_TILE_W) * elem_bytes,
                        size_t(w) * elem_bytes);
    }
}

struct RenderOut {
    uint8_t* rgb8; double* rgb_f64; double* final_state; int32_t* obj_id; int32_t* status; int32_t* nsteps;
};

// Is `p` host memory the GPUs can address directly (page-locked: cudaMallocHost / cudaHostAlloc /
// cudaHostRegister)?  Then the kernel reads and writes it in place over PCIe ("zero copy").
bool host_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// One implementation behind rtgr_render[_tiles|_resident] (rays generated on the device from `cam`)
// and rtgr_trace_canvas (rays read from the caller's Pixel array `px_host`, ni x nj, rgb written into
// it).  A page-locked Pixel array is used IN PLACE: the kernel loads pos/normal from it and stores rgb
// into it through the mapped address while it computes, so no copy brackets the kernel (64 B in and
// 24 B out per ray are spread over the whole kernel time, ~2 GB/s at 4K -- far below PCIe).  A
// pageable array is staged: uploaded, traced in its device copy, and copied back.
int render_impl(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                const rtgr_camera* cam, rtgr_pixel* px_host, int px_ni, int px_nj,
                int tile_offset, int tile_stride, const RenderOut& out,
                bool copy_back, rtgr_stats* stats) {
    if (!ctx) return fail("ctx is NULL");
    if (!cam && !px_host) return fail(px_ni || px_nj ? "pixels is NULL" : "camera is NULL");
    if (tile_stride < 1 || tile_offset < 0 || tile_offset >= tile_stride) return fail("bad tile_offset/tile_stride");
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, objs, n_objs, cam, sc, err)) return fail(err);
    rtgr_camera cam_px{};   // pixels mode: only ni, nj are meaningful
    if (px_host) {
        if (px_ni <= 0 || px_nj <= 0) return fail("canvas ni/nj must be positive");
        sc.ni = px_ni; sc.nj = px_nj;
        cam_px.ni = px_ni; cam_px.nj = px_nj;
        cam = &cam_px;
    }
    const double w0 = now_ms();
    const int variant = variant_of(params);
    const int64_t n = int64_t(cam->ni) * cam->nj;
    if (n >= (int64_t(1) << 31)) return fail("frames of 2^31 pixels or more are not supported");
    const int D = int(ctx->devs.size());
    const bool px_zero_copy = px_host && host_pinned(px_host);
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    // S tile shards: one per device (device k of D takes every D-th tile of the caller's selection), or ONE that
    // all devices draw from together (shared queue: head, inputs and outputs in device 0's memory)
    const bool shared_q = D > 1 && multi_queue_shared(ctx);
    const int S = shared_q ? 1 : D;
    struct Sel { int off, stride; int64_t count; int tiles_x; };
    std::vector<Sel> sel(S);
    for (int k = 0; k < S; ++k) {
        sel[k].off = tile_offset + k * tile_stride;
        sel[k].stride = tile_stride * S;
        rtgr::tile_selection(cam->ni, cam->nj, sel[k].off, sel[k].stride, sel[k].tiles_x, sel[k].count);
    }
    // Which tiles a shard gets, and in what order it works through them, are two decisions.
    //  * WHICH: in a Kerr-Schild scene the tiles are sorted by the impact parameter of their centre ray
    //    (a cost proxy: rays passing near the hole take 10-30x more steps, see tile_order_by_impact) and
    //    dealt round-robin over that sorted list, so that every rank / device gets the same cost mix
    //    (with plain index round-robin the shards of an 8K frame differed by 5-7 % in kernel time).
    //  * ORDER: each shard works through its tiles in the same impact order, expensive first, which leaves
    //    only cheap uniform rays for the end of the launch.  (Before the warps drew their rays in private
    //    patch-sized chunks, row-major order was ~4 % faster on big shares because it kept a warp's lanes
    //    alike; with the chunks impact order wins at every size: 335 vs 340 ms at 4K, 22.5 vs 24.4 ms at
    //    960x540, profiles/r01y_tile_order.log.)  RTGR_TILE_ORDER=row|impact|shuffle overrides the order.
    // Results never depend on either decision.
    std::vector<std::vector<int32_t>> lists(S);
    {
        const char* mode = getenv("RTGR_TILE_ORDER");
        bool impact_order = true;
        if (mode) impact_order = (mode[0] == 'i' || mode[0] == 's');
        std::vector<int32_t> sorted;
        if (params->metric == RTGR_KERR_SCHILD) {
            // The keys are recomputed on every call (one pixel per tile in canvas mode) and only the SORT is reused
            // when they are unchanged: which tiles a tile_offset/tile_stride shard owns follows from the sorted
            // list, so every rank must derive it from the exact, complete keys -- never from a sampled shortcut.
            std::vector<double> keys = px_host ? rtgr::tile_impact_keys_pixels(px_host, px_ni, px_nj) : rtgr::tile_impact_keys(*cam);
            if (keys != ctx->order_keys) {
                ctx->order_sorted = rtgr::tiles_sorted_by_key(keys);
                ctx->order_keys.swap(keys);
            }
            sorted = ctx->order_sorted;
        }
        if (mode && !sorted.empty() && mode[0] == 's') {   // deterministic shuffle: worst case, experiments only
            unsigned long long z = 88172645463325252ull;
            for (size_t i = sorted.size() - 1; i > 0; --i) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; std::swap(sorted[i], sorted[z % (i + 1)]); }
        }
        if (!sorted.empty()) {
            for (int k = 0; k < S; ++k) {
                lists[k].reserve(size_t(sel[k].count));
                for (int64_t m = 0; m < sel[k].count; ++m) lists[k].push_back(sorted[size_t(sel[k].off + m * sel[k].stride)]);
                if (!impact_order) std::sort(lists[k].begin(), lists[k].end());
            }
        }
    }
    for (int k = 0; k < D; ++k) {
        Device& d = ctx->devs[k];
        Device& b = shared_q ? ctx->devs[0] : d;   // whose memory holds this launch's inputs and outputs
        const int sh = shared_q ? 0 : k;           // which tile shard this device works on
        CU(cudaSetDevice(d.id));
        if (upload_scene(d, sc, um)) return -1;
        Job job{};
        job.mode = rtgr::JOB_RENDER;
        job.tiles_x = sel[sh].tiles_x; job.tile_offset = sel[sh].off; job.tile_stride = sel[sh].stride;
        job.queue_scope = shared_q ? 1 : 0;
        if (!lists[sh].empty()) {     // explicit tile list: ordinal m -> lists[sh][m] (every device keeps its own copy)
            if (upload_tile_list(d, lists[sh])) return -1;
            job.tile_order = (const int32_t*)d.order.p;
            job.tile_offset = 0; job.tile_stride = 1;
        }
        job.total = sel[sh].count * (RTGR_TILE_W * RTGR_TILE_H);
        job.rgb_stride = 3;
        if (px_host) {
            double* dpx = nullptr;
            if (px_zero_copy) {
                void* mapped = nullptr;
                CU(cudaHostGetDevicePointer(&mapped, px_host, 0));
                dpx = (double*)mapped;
            } else {
                if (ensure(b.pixels, size_t(n) * sizeof(rtgr_pixel))) return -1;
                b.resident_n = 0;
                if (&b == &d)     // staged once, into the memory of the device that owns the buffers
                    CU(cudaMemcpyAsync(b.pixels.p, px_host, size_t(n) * sizeof(rtgr_pixel), cudaMemcpyHostToDevice, d.stream));
                dpx = (double*)b.pixels.p;
            }
            job.pixels_in = dpx;
            job.rgb_f64 = dpx + 8;     // the rgb field of Pixel (src:446-450), written in place (src:532)
            job.rgb_stride = 11;
            // (trace_pixels_kernel: the colours go back as whole rows of a patch; wanted when four or more devices write
            // into one page-locked host canvas)
            job.stage_canvas = (px_zero_copy && ctx->devs.size() >= 4) ? 2 : 1;
        }
        // (k == 0 runs first, with device 0 current: in shared mode the buffers exist by the time k > 0 asks)
        if (out.rgb8 || (!copy_back && !px_host)) { if (ensure(b.rgb8, size_t(n) * 3)) return -1; job.rgb8 = (uint8_t*)b.rgb8.p; }
        if (out.rgb_f64) { if (ensure(b.rgbf, size_t(n) * 24)) return -1; job.rgb_f64 = (double*)b.rgbf.p; }
        if (out.final_state) { if (ensure(b.fstate, size_t(n) * 64)) return -1; job.final_state = (double*)b.fstate.p; }
        if (out.obj_id) { if (ensure(b.objid, size_t(n) * 4)) return -1; job.obj_id = (int32_t*)b.objid.p; }
        if (out.status) { if (ensure(b.status, size_t(n) * 4)) return -1; job.status = (int32_t*)b.status.p; }
        if (out.nsteps) { if (ensure(b.nsteps, size_t(n) * 4)) return -1; job.nsteps = (int32_t*)b.nsteps.p; }
        job.stage_rgb8 = stage_rgb8_wanted(job, cam ? cam->ni : px_ni, /*remote=*/&b != &d) ? 1 : 0;
        if (shared_q) {
            // device 0 zeroes the shared head (after its staging copy, in stream order) and signals; the others wait
            if (k == 0) {
                CU(cudaMemsetAsync(b.d_next, 0, sizeof(unsigned long long), d.stream));
                CU(cudaEventRecord(ctx->ev_shared, d.stream));
            } else {
                CU(cudaStreamWaitEvent(d.stream, ctx->ev_shared, 0));
            }
        }
        if (launch_trace(d, variant, job, um, shared_q ? b.d_next : nullptr)) return -1;
    }
    if (shared_q)   // the kernels of ALL devices have written into device 0's buffers
        for (auto& d : ctx->devs) { CU(cudaSetDevice(d.id)); CU(cudaStreamSynchronize(d.stream)); }
    if (copy_back) {
        const bool whole = (S == 1 && tile_stride == 1);
        struct Item { void* dst; DevBuf Device::*buf; size_t elem; };
        const Item items[] = {
            {(px_host && !px_zero_copy) ? (void*)px_host : nullptr, &Device::pixels, sizeof(rtgr_pixel)},
            {out.rgb8, &Device::rgb8, 3}, {out.rgb_f64, &Device::rgbf, 24}, {out.final_state, &Device::fstate, 64},
            {out.obj_id, &Device::objid, 4}, {out.status, &Device::status, 4}, {out.nsteps, &Device::nsteps, 4}};
        if (whole) {
            Device& d = ctx->devs[0];
            CU(cudaSetDevice(d.id));
            for (const Item& it : items)
                if (it.dst) CU(cudaMemcpyAsync(it.dst, (d.*(it.buf)).p, size_t(n) * it.elem, cudaMemcpyDeviceToHost, d.stream));
            CU(cudaStreamSynchronize(d.stream));
        } else {
            // each device packs ITS tiles of every requested buffer into a compact tile-major block, returns that
            // block into pinned staging, and a host thread per device scatters it into the caller's image (the
            // "final host gather"): 1/S of the frame crosses PCIe per device
            size_t total_elem = 0;
            for (const Item& it : items) if (it.dst) total_elem += it.elem;
            constexpr size_t TPX = size_t(RTGR_TILE_W) * RTGR_TILE_H;
            for (int k = 0; k < S; ++k) {      // shard k lives on device k (shared queue: the one shard on device 0)
                Device& d = ctx->devs[k];
                CU(cudaSetDevice(d.id));
                const size_t slots = size_t(sel[k].count) * TPX;
                if (slots == 0) continue;
                if (ensure(d.pack, slots * total_elem)) return -1;
                if (ensure_pinned(d.h_stage, slots * total_elem)) return -1;
                const int32_t* dlist = lists[k].empty() ? nullptr : (const int32_t*)d.order.p;
                const int blocks = int(std::min<int64_t>(sel[k].count, int64_t(d.sm_count) * 8));
                size_t off = 0;
                for (const Item& it : items)
                    if (it.dst) {
                        pack_tiles_kernel<<<blocks, 256, 0, d.stream>>>((const uint8_t*)(d.*(it.buf)).p, (uint8_t*)d.pack.p + off, dlist,
                                                                        sel[k].off, sel[k].stride, (long long)sel[k].count,
                                                                        sel[k].tiles_x, cam->ni, cam->nj, int(it.elem));
                        off += slots * it.elem;
                    }
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(d.h_stage.p, d.pack.p, slots * total_elem, cudaMemcpyDeviceToHost, d.stream));
            }
            std::vector<std::thread> th;
            for (int k = 0; k < S; ++k) {
                th.emplace_back([&, k]() {
                    Device& d = ctx->devs[k];
                    cudaSetDevice(d.id);
                    cudaStreamSynchronize(d.stream);
                    const size_t slots = size_t(sel[k].count) * TPX;
                    size_t off = 0;
                    for (const Item& it : items)
                        if (it.dst) {
                            scatter_packed_tiles((const uint8_t*)d.h_stage.p + off, (uint8_t*)it.dst, cam->ni, cam->nj, it.elem,
                                                 sel[k].tiles_x, sel[k].off, sel[k].stride, sel[k].count,
                                                 lists[k].empty() ? nullptr : lists[k].data());
                            off += slots * it.elem;
                        }
                });
            }
            for (auto& t : th) t.join();
        }
    }
    for (auto& d : ctx->devs) { CU(cudaSetDevice(d.id)); CU(cudaStreamSynchronize(d.stream)); }
    return collect_stats(ctx, stats, now_ms() - w0);
}

// Pixels mode over D devices: device k owns the 1024-ray blocks b with b % D == k.  Blocks are
// moved with 2-D copies (row = one block), so each device holds a compact local array.
constexpr int64_t PBLOCK = 1024;

struct PixelSplit {
    int64_t full_rows;   // number of complete 1024-ray blocks this device owns
    int64_t tail;        // rays in a trailing partial block (owned by exactly one device)
    int64_t tail_start;  // global index of the partial block
    int64_t local_n() const { return full_rows * PBLOCK + tail; }
};

PixelSplit split_pixels(int64_t n, int k, int D) {
    PixelSplit s{};
    const int64_t nfull = n / PBLOCK;
    s.full_rows = (nfull > k) ? (nfull - k + D - 1) / D : 0;
    const int64_t rem = n - nfull * PBLOCK;
    if (rem > 0 && (nfull % D) == k) { s.tail = rem; s.tail_start = nfull * PBLOCK; }
    return s;
}

int copy_blocks(Device& d, int k, int D, const PixelSplit& sp, void* dev, void* host, size_t elem, bool to_device) {
    const size_t row = size_t(PBLOCK) * elem;
    uint8_t* h = (uint8_t*)host + size_t(k) * row;
    if (sp.full_rows > 0) {
        if (to_device)
            CU(cudaMemcpy2DAsync(dev, row, h, row * D, row, size_t(sp.full_rows), cudaMemcpyHostToDevice, d.stream));
        else
            CU(cudaMemcpy2DAsync(h, row * D, dev, row, row, size_t(sp.full_rows), cudaMemcpyDeviceToHost, d.stream));
    }
    if (sp.tail > 0) {
        uint8_t* ht = (uint8_t*)host + size_t(sp.tail_start) * elem;
        uint8_t* dt = (uint8_t*)dev + size_t(sp.full_rows) * row;
        if (to_device) CU(cudaMemcpyAsync(dt, ht, size_t(sp.tail) * elem, cudaMemcpyHostToDevice, d.stream));
        else CU(cudaMemcpyAsync(ht, dt, size_t(sp.tail) * elem, cudaMemcpyDeviceToHost, d.stream));
    }
    return 0;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* rtgr_last_error(void) { return g_err.c_str(); }
int rtgr_version(void) { return RTGR_VERSION; }

void rtgr_default_params(rtgr_params* p, int metric) {
    if (!p) return;
    p->metric = metric; p->r_formula = RTGR_R_AS_WRITTEN;
    p->M = 1.0; p->a = 0.0;
    p->lambda0 = 0.0; p->lambda1 = 100.0;
    p->reltol = p->abstol = 1.8189894035458565e-12;  // eps(Float64)^(3/4)
    p->hit_threshold = 0.01;
    p->interp_points = 10; p->maxiters = 100000;
}

int rtgr_create(rtgr_ctx** out, const int* device_ids, int n_devices) {
    if (!out) return fail("ctx out pointer is NULL");
    *out = nullptr;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail == 0)
        return fail(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (n_devices <= 0) n_devices = avail;
    auto* ctx = new rtgr_ctx();
    for (int k = 0; k < n_devices; ++k) {
        Device d;
        d.id = device_ids ? device_ids[k] : k;
        if (d.id < 0 || d.id >= avail) { delete ctx; return fail("device id out of range"); }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d.id) != cudaSuccess) { delete ctx; return fail("cudaGetDeviceProperties failed"); }
        if (prop.major < 10) { delete ctx; return fail("device is not Blackwell (sm_100a required)"); }
        d.sm_count = prop.multiProcessorCount;
        std::memcpy(d.uuid, prop.uuid.bytes, 16);
        if (cudaSetDevice(d.id) != cudaSuccess || cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&d.ev0) != cudaSuccess || cudaEventCreate(&d.ev1) != cudaSuccess ||
            cudaMalloc(&d.d_next, sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&d.d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess) {
            delete ctx;
            return fail(std::string("device setup failed: ") + cudaGetErrorString(cudaGetLastError()));
        }
        ctx->devs.push_back(d);
    }
    *out = ctx;
    return 0;
}

void rtgr_frame_close(rtgr_frame* fr);

void rtgr_destroy(rtgr_ctx* ctx) {
    if (!ctx) return;
    while (!ctx->frames.empty()) rtgr_frame_close(ctx->frames.back());   // frames left open by the caller
    if (ctx->ev_shared) { cudaSetDevice(ctx->devs[0].id); cudaEventDestroy(ctx->ev_shared); }
    for (auto& m : ctx->metrics) if (m.alive && m.lib) cudaLibraryUnload(m.lib);
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.id);
        cudaStreamSynchronize(d.stream);
        for (DevBuf* b : {&d.pixels, &d.rgb8, &d.rgbf, &d.fstate, &d.objid, &d.status, &d.nsteps, &d.scratch, &d.order})
            if (b->p) cudaFree(b->p);
        if (d.h_stage.p) cudaFreeHost(d.h_stage.p);
        cudaFree(d.d_next); cudaFree(d.d_counters);
        cudaEventDestroy(d.ev0); cudaEventDestroy(d.ev1);
        for (cudaEvent_t e : d.chunk_ev) cudaEventDestroy(e);
        cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

int rtgr_device_count(const rtgr_ctx* ctx) { return ctx ? int(ctx->devs.size()) : 0; }

void* rtgr_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void rtgr_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int rtgr_render_tiles(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                      const rtgr_camera* cam, int tile_offset, int tile_stride, uint8_t* rgb8, double* rgb_f64,
                      double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps, rtgr_stats* stats) {
    RenderOut out{rgb8, rgb_f64, final_state, obj_id, status, nsteps};
    return render_impl(ctx, params, objs, n_objs, cam, nullptr, 0, 0, tile_offset, tile_stride, out, true, stats);
}

int rtgr_render(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                const rtgr_camera* cam, uint8_t* rgb8, double* rgb_f64, double* final_state, int32_t* obj_id,
                int32_t* status, int32_t* nsteps, rtgr_stats* stats) {
    return rtgr_render_tiles(ctx, params, objs, n_objs, cam, 0, 1, rgb8, rgb_f64, final_state, obj_id, status, nsteps,
                             stats);
}

int rtgr_render_resident(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                         const rtgr_camera* cam, int tile_offset, int tile_stride, rtgr_stats* stats) {
    RenderOut out{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    return render_impl(ctx, params, objs, n_objs, cam, nullptr, 0, 0, tile_offset, tile_stride, out, false, stats);
}

int rtgr_trace_canvas(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                      rtgr_pixel* pixels, int ni, int nj, int tile_offset, int tile_stride, double* final_state,
                      int32_t* obj_id, int32_t* status, int32_t* nsteps, rtgr_stats* stats) {
    if (!pixels) return fail("pixels is NULL");
    RenderOut out{nullptr, nullptr, final_state, obj_id, status, nsteps};
    return render_impl(ctx, params, objs, n_objs, nullptr, pixels, ni, nj, tile_offset, tile_stride, out, true, stats);
}

int rtgr_host_register(void* p, uint64_t bytes) {
    if (!p || !bytes) return fail("NULL argument");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}
int rtgr_host_unregister(void* p) {
    if (!p) return fail("NULL argument");
    CU(cudaHostUnregister(p));
    return 0;
}
int rtgr_host_is_pinned(const void* p) { return (p && host_pinned(p)) ? 1 : 0; }

// ---- user-supplied metrics ----------------------------------------------------
int rtgr_metric_check(const char* source, char* log, uint64_t log_capacity) {
    std::vector<char> cubin; std::string lg, err;
    const bool ok = rtgr_jit::compile(source, cubin, lg, err);
    if (log && log_capacity) {
        const std::string& msg = ok ? lg : err;
        const size_t n = std::min<size_t>(msg.size(), size_t(log_capacity) - 1);
        std::memcpy(log, msg.data(), n); log[n] = 0;
    }
    return ok ? 0 : fail(err);
}

int rtgr_metric_compile(rtgr_ctx* ctx, const char* source, int32_t* metric_id) {
    if (!ctx || !metric_id) return fail("NULL argument");
    *metric_id = -1;
    std::vector<char> cubin; std::string lg, err;
    if (!rtgr_jit::compile(source, cubin, lg, err)) return fail(err);
    UserMetric um;
    CU(cudaSetDevice(ctx->devs[0].id));
    CU(cudaLibraryLoadData(&um.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    CU(cudaLibraryGetKernel(&um.k_trace, um.lib, "rtgr_user_trace"));
    CU(cudaLibraryGetKernel(&um.k_trace_paths, um.lib, "rtgr_user_trace_paths"));
    CU(cudaLibraryGetKernel(&um.k_trace_stage, um.lib, "rtgr_user_trace_stage"));
    CU(cudaLibraryGetKernel(&um.k_rhs, um.lib, "rtgr_user_rhs"));
    CU(cudaLibraryGetKernel(&um.k_canvas, um.lib, "rtgr_user_canvas"));
    um.alive = true;
    ctx->metrics.push_back(um);
    *metric_id = RTGR_USER_METRIC_BASE + int32_t(ctx->metrics.size()) - 1;
    return 0;
}

int rtgr_metric_set_params(rtgr_ctx* ctx, int32_t metric_id, const double* par, int n) {
    rtgr_params p{}; p.metric = metric_id;
    UserMetric* um = nullptr;
    if (metric_id < RTGR_USER_METRIC_BASE || user_metric_of(ctx, &p, &um) || !um) return fail("unknown user metric id");
    if (n < 0 || n > 16 || (n > 0 && !par)) return fail("a user metric takes at most 16 parameters");
    std::memset(um->par, 0, sizeof(um->par));
    for (int i = 0; i < n; ++i) um->par[i] = par[i];
    return 0;
}

int rtgr_metric_release(rtgr_ctx* ctx, int32_t metric_id) {
    rtgr_params p{}; p.metric = metric_id;
    UserMetric* um = nullptr;
    if (metric_id < RTGR_USER_METRIC_BASE || user_metric_of(ctx, &p, &um) || !um) return fail("unknown user metric id");
    for (auto& d : ctx->devs) { cudaSetDevice(d.id); cudaStreamSynchronize(d.stream); }
    cudaLibraryUnload(um->lib);
    um->lib = nullptr; um->alive = false;
    return 0;
}

int rtgr_make_canvas(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_camera* cam, rtgr_pixel* pixels) {
    if (!ctx || !cam || !pixels) return fail("NULL argument");
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, nullptr, 0, cam, sc, err)) return fail(err);
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    Device& d = ctx->devs[0];
    CU(cudaSetDevice(d.id));
    const size_t bytes = size_t(cam->ni) * cam->nj * sizeof(rtgr_pixel);
    if (ensure(d.pixels, bytes)) return -1;
    d.resident_n = 0;
    if (upload_scene(d, sc, um)) return -1;
    dim3 grid((cam->ni + 127) / 128, cam->nj);
    if (um) {
        double* dpx = (double*)d.pixels.p;
        void* args[] = {&dpx};
        CU(cudaLaunchKernel((const void*)um->k_canvas, grid, dim3(128), args, 0, d.stream));
    } else {
        with_variant(variant_of(params), [&](auto M, auto R) {
            canvas_kernel<decltype(M)::value, decltype(R)::value><<<grid, 128, 0, d.stream>>>((double*)d.pixels.p);
            return 0;
        });
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pixels, d.pixels.p, bytes, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return 0;
}

int rtgr_rhs_batch(rtgr_ctx* ctx, const rtgr_params* params, const double* states, int64_t n, double* derivs) {
    if (!ctx || !states || !derivs) return fail("NULL argument");
    if (n <= 0) return 0;
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, nullptr, 0, nullptr, sc, err)) return fail(err);
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    Device& d = ctx->devs[0];
    CU(cudaSetDevice(d.id));
    if (ensure(d.scratch, size_t(n) * 128)) return -1;
    double* din = (double*)d.scratch.p;
    double* dout = din + 8 * n;
    if (upload_scene(d, sc, um)) return -1;
    CU(cudaMemcpyAsync(din, states, size_t(n) * 64, cudaMemcpyHostToDevice, d.stream));
    const int grid = int((n + 255) / 256);
    if (um) {
        long long nn = n;
        void* args[] = {&din, &nn, &dout};
        CU(cudaLaunchKernel((const void*)um->k_rhs, dim3(grid), dim3(256), args, 0, d.stream));
    } else {
        with_variant(variant_of(params), [&](auto M, auto R) {
            rhs_kernel<decltype(M)::value, decltype(R)::value><<<grid, 256, 0, d.stream>>>(din, n, dout);
            return 0;
        });
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(derivs, dout, size_t(n) * 64, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return 0;
}

static int trace_pixels_impl(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                             rtgr_pixel* pixels, int64_t n, bool upload, bool download, double* final_state,
                             int32_t* obj_id, int32_t* status, int32_t* nsteps, rtgr_stats* stats) {
    if (!ctx) return fail("ctx is NULL");
    if (n < 0) return fail("n is negative");
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, objs, n_objs, nullptr, sc, err)) return fail(err);
    if (n == 0) { if (stats) std::memset(stats, 0, sizeof(*stats)); return 0; }
    if ((upload || download) && !pixels) return fail("pixels is NULL");
    const double w0 = now_ms();
    const int variant = variant_of(params);
    const int D = int(ctx->devs.size());
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    // host threads (= D2H pieces) per device for the rgb write-back
    const int hw = int(std::thread::hardware_concurrency());
    const int nchunk = (n < 65536) ? 1 : std::max(1, std::min(8, (hw > 0 ? hw : 8) / D));
    std::vector<PixelSplit> sp(D);
    for (int k = 0; k < D; ++k) {
        Device& d = ctx->devs[k];
        sp[k] = split_pixels(n, k, D);
        const int64_t ln = sp[k].local_n();
        CU(cudaSetDevice(d.id));
        if (!upload && d.resident_n != ln) return fail("no resident pixel buffer of this size (call rtgr_upload_pixels)");
        if (ln == 0) { d.resident_n = 0; continue; }
        if (upload) {
            if (ensure(d.pixels, size_t(ln) * sizeof(rtgr_pixel))) return -1;
            if (copy_blocks(d, k, D, sp[k], d.pixels.p, pixels, sizeof(rtgr_pixel), true)) return -1;
            d.resident_n = ln;
        }
        if (upload_scene(d, sc, um)) return -1;
        Job job{};
        job.mode = rtgr::JOB_PIXELS;
        job.total = ln;
        job.pixels_in = (const double*)d.pixels.p;
        if (ensure(d.rgbf, size_t(ln) * 24)) return -1;
        job.rgb_f64 = (double*)d.rgbf.p;
        job.rgb_stride = 3;
        if (final_state) { if (ensure(d.fstate, size_t(ln) * 64)) return -1; job.final_state = (double*)d.fstate.p; }
        if (obj_id) { if (ensure(d.objid, size_t(ln) * 4)) return -1; job.obj_id = (int32_t*)d.objid.p; }
        if (status) { if (ensure(d.status, size_t(ln) * 4)) return -1; job.status = (int32_t*)d.status.p; }
        if (nsteps) { if (ensure(d.nsteps, size_t(ln) * 4)) return -1; job.nsteps = (int32_t*)d.nsteps.p; }
        if (launch_trace(d, variant, job, um)) return -1;
        if (download) {
            // rgb comes back compact (n x 3) into pinned staging, in `nchunk` pieces with an event after
            // each, so that the host threads which write the rgb field of the caller's Pixel array
            // (src:532) work on piece c while piece c+1 is still on the bus; the optional arrays go
            // straight to the caller.
            if (ensure_pinned(d.h_stage, size_t(ln) * 24)) return -1;
            while (int(d.chunk_ev.size()) < nchunk) {
                cudaEvent_t e;
                CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                d.chunk_ev.push_back(e);
            }
            for (int c = 0; c < nchunk; ++c) {
                const int64_t lo = ln * c / nchunk, hi = ln * (c + 1) / nchunk;
                if (hi > lo)
                    CU(cudaMemcpyAsync((uint8_t*)d.h_stage.p + size_t(lo) * 24, (uint8_t*)d.rgbf.p + size_t(lo) * 24,
                                       size_t(hi - lo) * 24, cudaMemcpyDeviceToHost, d.stream));
                CU(cudaEventRecord(d.chunk_ev[c], d.stream));
            }
            if (final_state && copy_blocks(d, k, D, sp[k], d.fstate.p, final_state, 64, false)) return -1;
            if (obj_id && copy_blocks(d, k, D, sp[k], d.objid.p, obj_id, 4, false)) return -1;
            if (status && copy_blocks(d, k, D, sp[k], d.status.p, status, 4, false)) return -1;
            if (nsteps && copy_blocks(d, k, D, sp[k], d.nsteps.p, nsteps, 4, false)) return -1;
        }
    }
    if (download) {
        std::vector<std::thread> th;
        for (int k = 0; k < D; ++k) {
            const int64_t ln = sp[k].local_n();
            if (ln == 0) continue;
            for (int c = 0; c < nchunk; ++c) {
                th.emplace_back([&, k, c, ln]() {
                    Device& d = ctx->devs[k];
                    cudaSetDevice(d.id);
                    cudaEventSynchronize(d.chunk_ev[c]);
                    const int64_t lo = ln * c / nchunk, hi = ln * (c + 1) / nchunk;
                    const int64_t nfull = sp[k].full_rows * PBLOCK;
                    const double* src = (const double*)d.h_stage.p + 3 * lo;
                    for (int64_t l = lo; l < hi; ++l, src += 3) {
                        // local ray l of device k -> global ray index (1024-ray blocks dealt round-robin)
                        const int64_t g = (l < nfull) ? ((l / PBLOCK) * D + k) * PBLOCK + (l % PBLOCK)
                                                      : sp[k].tail_start + (l - nfull);
                        pixels[g].rgb[0] = src[0]; pixels[g].rgb[1] = src[1]; pixels[g].rgb[2] = src[2];
                    }
                });
            }
        }
        for (auto& t : th) t.join();
    }
    for (auto& d : ctx->devs) { CU(cudaSetDevice(d.id)); CU(cudaStreamSynchronize(d.stream)); }
    return collect_stats(ctx, stats, now_ms() - w0);
}

// Ray paths: the reference's solve keeps every accepted step (save_everystep, appendix
// A) although trace_rays only reads sol[end]; this returns them.  Single device, n is expected to be
// small (the output is n x max_points x 72 bytes).
int rtgr_trace_paths(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                     const double* states0, int64_t n, int32_t max_points, double* paths, int32_t* npoints,
                     double* final_state, int32_t* obj_id, int32_t* status, rtgr_stats* stats) {
    if (!ctx) return fail("ctx is NULL");
    if (n < 0) return fail("n is negative");
    if (max_points < 2) return fail("max_points must be at least 2 (first and last point)");
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, objs, n_objs, nullptr, sc, err)) return fail(err);
    if (n == 0) { if (stats) std::memset(stats, 0, sizeof(*stats)); return 0; }
    if (!states0 || !paths || !npoints) return fail("NULL argument");
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    const double w0 = now_ms();
    Device& d = ctx->devs[0];
    CU(cudaSetDevice(d.id));
    std::vector<double> px(size_t(n) * 11, 0.0);      // the kernel reads rays as Pixel records (pos, normal, rgb)
    for (int64_t i = 0; i < n; ++i) std::memcpy(&px[size_t(i) * 11], states0 + 8 * i, 64);
    const size_t pbytes = size_t(n) * size_t(max_points) * 72;
    if (ensure(d.pixels, px.size() * 8) || ensure(d.scratch, pbytes) || ensure(d.nsteps, size_t(n) * 4) ||
        ensure(d.rgbf, size_t(n) * 24)) return -1;
    d.resident_n = 0;
    CU(cudaMemcpyAsync(d.pixels.p, px.data(), px.size() * 8, cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemsetAsync(d.scratch.p, 0, pbytes, d.stream));
    if (upload_scene(d, sc, um)) return -1;
    Job job{};
    job.mode = rtgr::JOB_PIXELS; job.total = n;
    job.pixels_in = (const double*)d.pixels.p;
    job.rgb_f64 = (double*)d.rgbf.p; job.rgb_stride = 3;
    job.paths = (double*)d.scratch.p; job.npoints = (int32_t*)d.nsteps.p; job.max_points = max_points;
    if (final_state) { if (ensure(d.fstate, size_t(n) * 64)) return -1; job.final_state = (double*)d.fstate.p; }
    if (obj_id) { if (ensure(d.objid, size_t(n) * 4)) return -1; job.obj_id = (int32_t*)d.objid.p; }
    if (status) { if (ensure(d.status, size_t(n) * 4)) return -1; job.status = (int32_t*)d.status.p; }
    if (launch_trace(d, variant_of(params), job, um)) return -1;
    CU(cudaMemcpyAsync(paths, d.scratch.p, pbytes, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaMemcpyAsync(npoints, d.nsteps.p, size_t(n) * 4, cudaMemcpyDeviceToHost, d.stream));
    if (final_state) CU(cudaMemcpyAsync(final_state, d.fstate.p, size_t(n) * 64, cudaMemcpyDeviceToHost, d.stream));
    if (obj_id) CU(cudaMemcpyAsync(obj_id, d.objid.p, size_t(n) * 4, cudaMemcpyDeviceToHost, d.stream));
    if (status) CU(cudaMemcpyAsync(status, d.status.p, size_t(n) * 4, cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return collect_stats(ctx, stats, now_ms() - w0);
}

int rtgr_trace_pixels(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                      rtgr_pixel* pixels, int64_t n, double* final_state, int32_t* obj_id, int32_t* status,
                      int32_t* nsteps, rtgr_stats* stats) {
    return trace_pixels_impl(ctx, params, objs, n_objs, pixels, n, true, true, final_state, obj_id, status, nsteps, stats);
}

int rtgr_upload_pixels(rtgr_ctx* ctx, const rtgr_pixel* pixels, int64_t n) {
    if (!ctx || !pixels) return fail("NULL argument");
    const int D = int(ctx->devs.size());
    for (int k = 0; k < D; ++k) {
        Device& d = ctx->devs[k];
        PixelSplit sp = split_pixels(n, k, D);
        CU(cudaSetDevice(d.id));
        d.resident_n = sp.local_n();
        if (d.resident_n == 0) continue;
        if (ensure(d.pixels, size_t(d.resident_n) * sizeof(rtgr_pixel))) return -1;
        if (copy_blocks(d, k, D, sp, d.pixels.p, (void*)pixels, sizeof(rtgr_pixel), true)) return -1;
    }
    for (auto& d : ctx->devs) { CU(cudaSetDevice(d.id)); CU(cudaStreamSynchronize(d.stream)); }
    return 0;
}

int rtgr_trace_resident(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                        rtgr_stats* stats) {
    if (!ctx) return fail("ctx is NULL");
    int64_t n = 0;
    const int D = int(ctx->devs.size());
    // recover n from the per-device resident sizes
    for (auto& d : ctx->devs) n += d.resident_n;
    if (n == 0) return fail("no resident pixel buffer (call rtgr_upload_pixels)");
    (void)D;
    return trace_pixels_impl(ctx, params, objs, n_objs, nullptr, n, false, false, nullptr, nullptr, nullptr, nullptr, stats);
}

// ---- cross-GPU dynamic tile queue over peer memory (north star item 4) -------------
int rtgr_frame_create(rtgr_ctx* ctx, int ni, int nj, rtgr_frame** out, uint8_t* ipc_handle) {
    if (!ctx || !out) return fail("NULL argument");
    *out = nullptr;
    if (ni <= 0 || nj <= 0) return fail("frame ni/nj must be positive");
    if (int64_t(ni) * nj >= (int64_t(1) << 31)) return fail("frames of 2^31 pixels or more are not supported");
    Device& d = ctx->devs[0];
    CU(cudaSetDevice(d.id));
    size_t bytes = FRAME_HEADER + size_t(ni) * size_t(nj) * 3;
    bytes = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);   // whole 2 MB pages: one allocation = one IPC mapping
    uint8_t* base = nullptr;
    CU(cudaMalloc(&base, bytes));
    const uint32_t info[4] = {FRAME_MAGIC, uint32_t(RTGR_VERSION), uint32_t(ni), uint32_t(nj)};
    cudaError_t e = cudaMemset(base, 0, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(base + FRAME_INFO, info, sizeof(info), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(base + FRAME_INFO + sizeof(info), d.uuid, 16, cudaMemcpyHostToDevice);   // bytes 80..95
    if (e != cudaSuccess) {
        cudaFree(base);
        return fail(std::string("rtgr_frame_create: ") + cudaGetErrorString(e));
    }
    if (ipc_handle) {
        // A driver/sandbox without IPC support leaves the handle all zero (rtgr_frame_open rejects it);
        // the frame is still good for the devices of this context.
        cudaIpcMemHandle_t h;
        std::memset(ipc_handle, 0, RTGR_IPC_HANDLE_BYTES);
        const cudaError_t ei = cudaIpcGetMemHandle(&h, base);
        if (ei == cudaSuccess) std::memcpy(ipc_handle, &h, sizeof(h));
        else { cudaGetLastError(); g_err = std::string("rtgr_frame_create: no IPC handle (in-process use only): ") + cudaGetErrorString(ei); }
    }
    auto* fr = new rtgr_frame();
    fr->ctx = ctx; fr->ni = ni; fr->nj = nj; fr->owner = true; fr->base = base; fr->home = d.id;
    std::memcpy(fr->owner_uuid, d.uuid, 16);
    ctx->frames.push_back(fr);
    *out = fr;
    return 0;
}

int rtgr_frame_open(rtgr_ctx* ctx, const uint8_t* ipc_handle, int ni, int nj, rtgr_frame** out) {
    if (!ctx || !out || !ipc_handle) return fail("NULL argument");
    *out = nullptr;
    Device& d = ctx->devs[0];
    CU(cudaSetDevice(d.id));
    bool zero = true;
    for (int i = 0; i < RTGR_IPC_HANDLE_BYTES; ++i) zero = zero && ipc_handle[i] == 0;
    if (zero) return fail("rtgr_frame_open: empty IPC handle (the owner could not export the frame)");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    uint32_t info[4] = {0, 0, 0, 0};
    const cudaError_t e = cudaMemcpy(info, (uint8_t*)p + FRAME_INFO, sizeof(info), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || info[0] != FRAME_MAGIC || info[1] != uint32_t(RTGR_VERSION) || int(info[2]) != ni || int(info[3]) != nj) {
        cudaIpcCloseMemHandle(p);
        if (e != cudaSuccess) return fail(std::string("rtgr_frame_open: ") + cudaGetErrorString(e));
        return fail(info[0] != FRAME_MAGIC ? "rtgr_frame_open: the handle does not refer to an rtgr_frame"
                                           : "rtgr_frame_open: frame size or library version differs from the owner's");
    }
    auto* fr = new rtgr_frame();
    fr->ctx = ctx; fr->ni = ni; fr->nj = nj; fr->owner = false; fr->base = (uint8_t*)p;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) fr->home = at.device;
    else cudaGetLastError();
    // (`home` is the ordinal the mapping is attached to in THIS process -- for an IPC mapping that is the opening
    // device, whichever GPU holds the memory; the owner's UUID in the header says whose memory it really is)
    if (cudaMemcpy(fr->owner_uuid, (uint8_t*)p + FRAME_INFO + sizeof(info), 16, cudaMemcpyDeviceToHost) != cudaSuccess) cudaGetLastError();
    // fail HERE (where the caller can still fall back to rtgr_render_tiles on all ranks together), not in the
    // middle of a frame, if this GPU cannot do atomics on the owner's memory
    if (enable_peer(d.id, fr->home, false)) {
        cudaIpcCloseMemHandle(p);
        delete fr;
        return -1;
    }
    ctx->frames.push_back(fr);
    *out = fr;
    return 0;
}

// One frame through the shared queue: rays from the camera (image into the frame's RGB8 buffer) or from a Pixel
// canvas in page-locked host memory that every participant has mapped (rgb written into it in place).
static int frame_impl(rtgr_frame* fr, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                      const rtgr_camera* cam, rtgr_pixel* px_host, int px_ni, int px_nj, rtgr_stats* stats) {
    if (!fr || !fr->ctx) return fail("frame is NULL");
    if (!cam && !px_host) return fail(px_ni || px_nj ? "pixels is NULL" : "camera is NULL");
    const int ni = cam ? cam->ni : px_ni, nj = cam ? cam->nj : px_nj;
    if (ni != fr->ni || nj != fr->nj) return fail(cam ? "the camera's ni/nj differ from the frame's" : "the canvas's ni/nj differ from the frame's");
    rtgr_ctx* ctx = fr->ctx;
    SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(params, objs, n_objs, cam, sc, err)) return fail(err);
    double* dpx = nullptr;
    if (px_host) {
        if (!host_pinned(px_host))
            return fail("rtgr_trace_canvas_frame: the canvas must be page-locked in this process (rtgr_alloc_pinned, or "
                        "rtgr_host_register on the mapping of the shared array)");
        void* mapped = nullptr;
        CU(cudaHostGetDevicePointer(&mapped, px_host, 0));
        dpx = (double*)mapped;
        sc.ni = ni; sc.nj = nj;
    }
    const double w0 = now_ms();
    const int variant = variant_of(params);
    UserMetric* um = nullptr;
    if (user_metric_of(ctx, params, &um)) return -1;
    int tiles_x = 0; int64_t ntiles = 0;
    rtgr::tile_selection(ni, nj, 0, 1, tiles_x, ntiles);
    // Queue order = the whole frame's tiles, expensive first: a pure function of the camera (or of the canvas's
    // contents, which all participants share), so every participant derives the same ordinal -> tile map
    // without talking to the others.
    const std::vector<int32_t>* order = nullptr;
    if (params->metric == RTGR_KERR_SCHILD) {
        std::vector<double> keys = px_host ? rtgr::tile_impact_keys_pixels(px_host, ni, nj) : rtgr::tile_impact_keys(*cam);
        if (keys != ctx->order_keys) {
            ctx->order_sorted = rtgr::tiles_sorted_by_key(keys);
            ctx->order_keys.swap(keys);
        }
        order = &ctx->order_sorted;
    }
    // Frames alternate between two queue heads.  The owner zeroes the head of the NEXT frame while this
    // one is being rendered (nobody touches it now: the participants are separated from the previous
    // frame, which used it, by the caller's barrier), so no reset-and-barrier step precedes a frame.
    auto* head_cur = (unsigned long long*)(fr->base + FRAME_HEAD_STRIDE * (fr->epoch & 1ull));
    auto* head_next = (unsigned long long*)(fr->base + FRAME_HEAD_STRIDE * ((fr->epoch + 1ull) & 1ull));
    for (size_t k = 0; k < ctx->devs.size(); ++k) {
        Device& d = ctx->devs[k];
        if (enable_peer(d.id, fr->home, fr->owner || k > 0)) return -1;
        CU(cudaSetDevice(d.id));
        if (upload_scene(d, sc, um)) return -1;
        if (fr->owner && k == 0) CU(cudaMemsetAsync(head_next, 0, sizeof(unsigned long long), d.stream));
        Job job{};
        job.mode = rtgr::JOB_RENDER;
        job.tiles_x = tiles_x; job.tile_offset = 0; job.tile_stride = 1;
        job.queue_scope = 1;
        if (order && upload_tile_list(d, *order)) return -1;
        if (order) job.tile_order = (const int32_t*)d.order.p;
        job.total = ntiles * (RTGR_TILE_W * RTGR_TILE_H);
        job.rgb_stride = 3;
        if (px_host) {
            job.pixels_in = dpx;
            job.rgb_f64 = dpx + 8;     // the rgb field of Pixel (src:446-450), written in place (src:532)
            job.rgb_stride = 11;
            // whole patches go back when four or more GPUs write into the one host canvas (rtgr_frame_set_participants;
            // the devices of this context count as that many)
            job.stage_canvas = (std::max(fr->participants, int(ctx->devs.size())) >= 4) ? 2 : 1;
        } else {
            job.rgb8 = fr->base + FRAME_HEADER;
        }
        job.stage_rgb8 = stage_rgb8_wanted(job, ni, /*remote=*/std::memcmp(d.uuid, fr->owner_uuid, 16) != 0) ? 1 : 0;
        if (launch_trace(d, variant, job, um, head_cur)) return -1;
    }
    for (auto& d : ctx->devs) { CU(cudaSetDevice(d.id)); CU(cudaStreamSynchronize(d.stream)); }
    fr->epoch += 1;
    return collect_stats(ctx, stats, now_ms() - w0);
}

int rtgr_render_frame(rtgr_frame* fr, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                      const rtgr_camera* cam, rtgr_stats* stats) {
    if (!fr || !fr->ctx) return fail("frame is NULL");
    if (!cam) return fail("camera is NULL");
    return frame_impl(fr, params, objs, n_objs, cam, nullptr, 0, 0, stats);
}

int rtgr_trace_canvas_frame(rtgr_frame* fr, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                            rtgr_pixel* pixels, int ni, int nj, rtgr_stats* stats) {
    if (!fr || !fr->ctx) return fail("frame is NULL");
    if (!pixels) return fail("pixels is NULL");
    return frame_impl(fr, params, objs, n_objs, nullptr, pixels, ni, nj, stats);
}

int rtgr_frame_set_participants(rtgr_frame* fr, int n) {
    if (!fr || !fr->ctx) return fail("frame is NULL");
    if (n < 0) return fail("the number of participants cannot be negative");
    fr->participants = n;
    return 0;
}

int rtgr_frame_read(rtgr_frame* fr, uint8_t* rgb8) {
    if (!fr || !rgb8) return fail("NULL argument");
    CU(cudaSetDevice(fr->ctx->devs[0].id));
    CU(cudaMemcpy(rgb8, fr->base + FRAME_HEADER, size_t(fr->ni) * size_t(fr->nj) * 3, cudaMemcpyDeviceToHost));
    return 0;
}

int rtgr_frame_clear(rtgr_frame* fr) {
    if (!fr) return fail("NULL argument");
    CU(cudaSetDevice(fr->ctx->devs[0].id));
    CU(cudaMemset(fr->base + FRAME_HEADER, 0, size_t(fr->ni) * size_t(fr->nj) * 3));
    CU(cudaDeviceSynchronize());
    return 0;
}

void rtgr_frame_close(rtgr_frame* fr) {
    if (!fr) return;
    auto& open = fr->ctx->frames;
    open.erase(std::remove(open.begin(), open.end(), fr), open.end());
    cudaSetDevice(fr->ctx->devs[0].id);
    if (fr->owner) cudaFree(fr->base);
    else cudaIpcCloseMemHandle(fr->base);
    delete fr;
}

int rtgr_fp64_peak(rtgr_ctx* ctx, int dev_index, double* tflops, double* sm_clock_mhz) {
    return rtgr_fp64_microbench(ctx, dev_index, 1, tflops, sm_clock_mhz);
}

int rtgr_fp64_microbench(rtgr_ctx* ctx, int dev_index, int mode, double* tflops, double* sm_clock_mhz) {
    if (!ctx || dev_index < 0 || dev_index >= int(ctx->devs.size())) return fail("bad device index");
    Device& d = ctx->devs[dev_index];
    CU(cudaSetDevice(d.id));
    int threads = 256, blocks = d.sm_count * 8, iters = 4096;
    int chains = 0;
    if (mode >= 100) {   // 100 + 10*log2(chains) + warps per scheduler: latency / parallelism probe, one block per SM
        chains = 1 << ((mode - 100) / 10);
        const int wps = (mode - 100) % 10;
        if (chains > 8 || wps < 1 || wps > 8) return fail("bad microbenchmark mode");
        threads = 128 * wps; blocks = d.sm_count; iters = 2048;
    }
    if (ensure(d.scratch, size_t(threads) * blocks * 8)) return -1;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(d.ev0, d.stream));
        double* o = (double*)d.scratch.p;
        const double seed = 1.0 + rep;
        switch (mode) {
            case 2: fp64_mix_kernel<2><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 3: fp64_mix_kernel<3><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 4: fp64_mix_kernel<4><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 5: fp64_mix_kernel<5><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 6: fp64_mix_kernel<6><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 7: fp64_mix_kernel<7><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 8: fp64_mix_kernel<8><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 9: fp64_mix_kernel<9><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 10: fp64_int_mix_kernel<10><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 11: fp64_int_mix_kernel<11><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            case 12: fp64_int_mix_kernel<12><<<blocks, threads, 0, d.stream>>>(o, iters, seed); break;
            default:
                if (chains == 1) fp64_chain_kernel<1><<<blocks, threads, 0, d.stream>>>(o, iters, seed);
                else if (chains == 2) fp64_chain_kernel<2><<<blocks, threads, 0, d.stream>>>(o, iters, seed);
                else if (chains == 4) fp64_chain_kernel<4><<<blocks, threads, 0, d.stream>>>(o, iters, seed);
                else if (chains == 8) fp64_chain_kernel<8><<<blocks, threads, 0, d.stream>>>(o, iters, seed);
                else fp64_peak_kernel<<<blocks, threads, 0, d.stream>>>(o, iters, seed);
                break;
        }
        CU(cudaEventRecord(d.ev1, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    const double flops = double(threads) * blocks * double(iters) * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d.id);
        *sm_clock_mhz = khz / 1000.0;
    }
    return 0;
}

}  // extern "C"
