/*
 * libraytracegr_cuda -- C ABI of the B200 (sm_100a) geodesic ray tracer.
 *
 * This is the drop-in boundary for ONE hot path of eschnett/RayTraceGR.jl:
 *
 *     trace_rays(metric, objs::Vector{Object{T}}, c::Canvas{T})::Canvas{T}
 *         reference: src/RayTraceGR.jl:482-536 (called from example1 :560, example2 :596)
 *
 * and everything it calls (geodesic :358-370, christoffel :321-331, dmetric :302-313,
 * kerr_schild :274-294, minkowski :262-264, the Tsit5 adaptive solve with a terminating
 * ContinuousCallback :488-511, classification + colouring :513-533, distance/objcolor
 * :399-428, min_distance :433-441).  The reference has no FFI of its own (it is pure
 * Julia), so the seam is that Julia function; INTEGRATION.md shows the `ccall` binding a
 * maintainer adds on the Julia side.
 *
 * Rules of the ABI
 *   - plain C, no C++ types, no exceptions cross the boundary;
 *   - every entry point returns int: 0 = OK, <0 = error; the message is available from
 *     rtgr_last_error() (thread-local);
 *   - the caller owns every host buffer; the library owns device memory inside the
 *     opaque context; calls on one context must be serialised by the caller;
 *   - there is NO CPU fallback: without a usable CUDA device rtgr_create fails.
 *
 * All floating point is IEEE binary64.  Index conventions follow the reference:
 * coordinates (t,x,y,z) = x^0..x^3, a ray state is the 8-vector (x^0..x^3,u^0..u^3)
 * (r2s, src:345-347), the canvas is column-major pixels[i,j] -> linear i + j*ni
 * (0-based), object ids are 1-based positions in the object list, 0 = "hit nothing"
 * (src:518-528).
 */
#ifndef RAYTRACEGR_CUDA_H
#define RAYTRACEGR_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTGR_VERSION 100
#define RTGR_MAX_OBJECTS 16

/* metric(x): the two built-in metrics of the reference (src:262, src:274). */
typedef enum { RTGR_MINKOWSKI = 0, RTGR_KERR_SCHILD = 1 } rtgr_metric_kind;

/* Kerr-Schild radius formula.  AS_WRITTEN is src:284 taken literally,
 *   r = sqrt(rho^2 - a^2)/2 + sqrt(a^2 z^2 + ((rho^2 - a^2)/2)^2),
 * which is what the reference's golden image sphere2.png was rendered with (parity
 * default).  CORRECTED is the textbook radius
 *   r = sqrt((rho^2 - a^2)/2 + sqrt(a^2 z^2 + ((rho^2 - a^2)/2)^2)). */
typedef enum { RTGR_R_AS_WRITTEN = 0, RTGR_R_CORRECTED = 1 } rtgr_r_formula;

/* Plane{T} (src:394-397) and Sphere{T} (src:409-413). */
typedef enum { RTGR_PLANE = 0, RTGR_SPHERE = 1 } rtgr_obj_kind;

/* Flat tagged form of one element of the reference's Vector{Object{T}}.
 * PLANE uses `time` only; SPHERE uses pos, vel (stored, never read: src:411) and radius
 * (negative radius = inside-out sky sphere, src:417). */
typedef struct {
    int32_t kind;
    int32_t _pad;
    double time;
    double pos[4];
    double vel[4];
    double radius;
} rtgr_object;

/* Everything the reference hard-codes, with the reference's values as defaults
 * (rtgr_default_params fills them in). */
typedef struct {
    int32_t metric;       /* rtgr_metric_kind                                            */
    int32_t r_formula;    /* rtgr_r_formula                                              */
    double M, a;          /* Kerr-Schild mass and spin; reference: 1, 0 (src:275-276)    */
    double lambda0;       /* affine-parameter span; reference (0, 100) (src:497)         */
    double lambda1;
    double reltol;        /* reference: eps(Float64)^(3/4) = 1.8189894035458565e-12      */
    double abstol;        /*            for both (src:485, :511)                         */
    double hit_threshold; /* colouring: object counts as hit if distance < 0.01 (src:519) */
    int32_t interp_points;/* ContinuousCallback dense-output sample points, 10           */
    int32_t maxiters;     /* solver step-attempt limit, 100000                           */
} rtgr_params;

/* Arguments of make_canvas(metric,pos,widthx,widthy,normal,ni,nj) (src:458-462). */
typedef struct {
    double pos[4];
    double widthx[4];
    double widthy[4];
    double normal[4];
    int32_t ni, nj;
} rtgr_camera;

/* Pixel{Float64} (src:446-450) is an isbits struct of 11 doubles, so a Julia
 * Array{Pixel{Float64},2} is one contiguous buffer of these. */
typedef struct {
    double pos[4];
    double normal[4];
    double rgb[3];
} rtgr_pixel;

/* Per-ray termination status (the reference ignores solver return codes, src:502-505;
 * the ray's last state is coloured whatever happened -- so do we). */
typedef enum {
    RTGR_STATUS_EVENT = 0,      /* stopped on an object (ContinuousCallback fired)         */
    RTGR_STATUS_LAMBDA_END = 1, /* reached lambda1 without touching anything               */
    RTGR_STATUS_MAXITERS = 2,   /* more than maxiters step attempts                        */
    RTGR_STATUS_DT_MIN = 3,     /* step size underflow                                     */
    RTGR_STATUS_NONFINITE = 4   /* NaN/Inf in the state (includes rho < a for AS_WRITTEN)  */
} rtgr_status;

/* Work counters of one call (summed over devices).  rhs_evals counts geodesic RHS
 * evaluations the kernel executed: 6 per step attempt + 2 per ray (the reference does
 * one more f(u0) at start whose value equals the first, and one after the event whose
 * value is never used). */
typedef struct {
    uint64_t rays;
    uint64_t rhs_evals;
    uint64_t steps_accepted;
    uint64_t steps_rejected;
    double kernel_ms; /* device time of the trace kernel(s), CUDA events; max over devices */
    double total_ms;  /* host wall time of the whole call, copies included                 */
    double drain_ms;  /* diagnostic: from the moment the ray queue ran empty to the end of the
                         kernel (the tail during which lanes idle); max over devices          */
} rtgr_stats;

typedef struct rtgr_ctx rtgr_ctx;

/* ---- life cycle -------------------------------------------------------------------- */

/* Create a context on the given CUDA devices (device_ids == NULL: devices 0..n-1;
 * n_devices <= 0: all visible devices). */
int rtgr_create(rtgr_ctx** ctx, const int* device_ids, int n_devices);
void rtgr_destroy(rtgr_ctx* ctx);
const char* rtgr_last_error(void);
int rtgr_version(void);
int rtgr_device_count(const rtgr_ctx* ctx);

/* Reference defaults for `metric` (src:275-276, :485, :497, :519). */
void rtgr_default_params(rtgr_params* p, int metric);

/* Pinned host memory helpers (page-locked buffers make the H2D/D2H legs run at full PCIe
 * rate; ordinary pageable memory is accepted everywhere too). */
void* rtgr_alloc_pinned(uint64_t bytes);
void rtgr_free_pinned(void* p);

/* ---- the hot path ------------------------------------------------------------------ */

/* Drop-in for trace_rays (src:483-536).  `pixels` is the reference's
 * Array{Pixel{Float64},2} (n elements of rtgr_pixel): pos and normal are read (the ray's
 * initial position and null 4-velocity, src:492-496) and rgb is written (src:527-532).
 * Optional outputs (NULL to skip): final_state n x 8 (= s2r(sol[end]), src:504),
 * obj_id n (omin of src:518-526), status n, nsteps n (accepted steps). */
int rtgr_trace_pixels(rtgr_ctx* ctx, const rtgr_params* params,
                      const rtgr_object* objs, int n_objs,
                      rtgr_pixel* pixels, int64_t n,
                      double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                      rtgr_stats* stats);

/* The same drop-in with the canvas SHAPE known (trace_rays receives a Canvas whose pixels are an
 * ni x nj column-major array, src:453-455, :483): rays are scheduled in 32x32-pixel tiles cut into
 * 8x4-pixel warp patches like rtgr_render, which keeps rays of similar cost together (~3 % faster
 * than the 1-D order of rtgr_trace_pixels), and only shard tile_offset of tile_stride (the partition
 * of rtgr_render_tiles; 0, 1 = whole canvas) is traced, so that several processes can share one canvas.  pos/normal of
 * the selected pixels are read, their rgb is written IN PLACE (src:527-532); the optional outputs
 * are in canvas order (index i + j*ni), untouched outside the selection.
 *
 * Zero copy: when `pixels` is page-locked host memory (rtgr_alloc_pinned, or any buffer -- e.g. the
 * memory of a Julia Array{Pixel{Float64},2} -- passed once to rtgr_host_register) the kernel reads
 * and writes it in place over PCIe while it computes; no H2D/D2H copy brackets the kernel.  Pageable
 * memory is accepted and staged through device copies (slower). */
int rtgr_trace_canvas(rtgr_ctx* ctx, const rtgr_params* params,
                      const rtgr_object* objs, int n_objs,
                      rtgr_pixel* pixels, int ni, int nj, int tile_offset, int tile_stride,
                      double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                      rtgr_stats* stats);

/* Page-lock (and map for all devices) an existing host buffer in place / undo it / query it. */
int rtgr_host_register(void* p, uint64_t bytes);
int rtgr_host_unregister(void* p);
int rtgr_host_is_pinned(const void* p);

/* Fused production path: make_canvas (src:458-478) runs on the device, then the same
 * trace; outputs are compact.  rgb8 is nj x ni x 3, row-major, row = j, col = i -- the
 * layout of the PNG the reference's example1/2 save (transpose at src:566-569, 8-bit value
 * = round(255*x)).  rgb_f64 is n x 3 in canvas order (i + j*ni).  Any output may be NULL.
 * With several devices in the context the screen is cut into 32x32-pixel tiles dealt round-robin
 * (in expensive-first order) to the devices; each device hands its tiles' rays to its warps
 * through a dynamic atomic queue.  Results do not depend on the device count. */
int rtgr_render(rtgr_ctx* ctx, const rtgr_params* params,
                const rtgr_object* objs, int n_objs, const rtgr_camera* cam,
                uint8_t* rgb8, double* rgb_f64,
                double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                rtgr_stats* stats);

/* As rtgr_render, but only shard `tile_offset` of a fixed partition of the screen's tiles into
 * `tile_stride` shards is traced (tile = RTGR_TILE_W x RTGR_TILE_H pixels; the tiles, sorted by
 * estimated cost in a Kerr-Schild scene and by index otherwise, are dealt round-robin to the
 * shards, so that all shards cost the same); pixels outside the shard are left untouched in the
 * output buffers.  This is how one process per GPU shares a frame: rank r of N passes (r, N); the N
 * shards cover every pixel exactly once. */
#define RTGR_TILE_W 32
#define RTGR_TILE_H 32
int rtgr_render_tiles(rtgr_ctx* ctx, const rtgr_params* params,
                      const rtgr_object* objs, int n_objs, const rtgr_camera* cam,
                      int tile_offset, int tile_stride,
                      uint8_t* rgb8, double* rgb_f64,
                      double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                      rtgr_stats* stats);

/* ---- one frame shared by several GPUs: the cross-GPU dynamic tile queue ------------------------ */

/* rtgr_render_tiles gives every shard a FIXED tile set.  A shared frame instead lets all GPUs draw
 * their rays from ONE queue as they fall idle: the frame is a single allocation in the owner GPU's
 * memory that holds the queue head and the RGB8 image; the other GPUs reach it as peer memory over
 * NVLink/NVSwitch -- further devices of the owner's context directly, other processes through a
 * CUDA IPC mapping.  Every participating kernel takes 8x4-pixel patches of the frame's 32x32-pixel
 * tiles from the shared head with one system-scope atomicAdd per patch (tiles in expensive-first
 * order, a pure function of the camera) and stores its pixels straight into the owner's image, so
 * the "final gather" of the reference's EnsembleThreads output (src:510-511, :513-533) is fused into
 * the trace kernel and the load balances itself whatever the cost distribution of the scene is.
 * There is no collective and no host copy between the GPUs.
 *
 * Protocol (one frame after the other):
 *   owner:  rtgr_frame_create(ctx, ni, nj, &frame, handle)   -- pass `handle` (64 bytes) to the others
 *   others: rtgr_frame_open(ctx, handle, ni, nj, &frame)      -- in other PROCESSES (one per GPU)
 *   every participant, once per frame, with the same params/objs/cam:  rtgr_render_frame(frame, ...)
 *   a barrier of the caller's (MPI, torch.distributed, ...) between consecutive frames and before
 *   rtgr_frame_read: a participant returns from rtgr_render_frame when the queue is empty and its own
 *   rays are done, which says nothing about the others.
 * The owner must take part in every frame (it re-arms the queue).  stats are this participant's share;
 * the shares of all participants add up to the frame (rays = ni*nj).  The image does not depend on who
 * traced what.  rtgr_destroy closes the frames of its context that are still open
 * (their handles are dead afterwards). */
typedef struct rtgr_frame rtgr_frame;
#define RTGR_IPC_HANDLE_BYTES 64
int rtgr_frame_create(rtgr_ctx* ctx, int ni, int nj, rtgr_frame** frame,
                      uint8_t* ipc_handle /* RTGR_IPC_HANDLE_BYTES out, or NULL: in-process use only */);
int rtgr_frame_open(rtgr_ctx* ctx, const uint8_t* ipc_handle, int ni, int nj, rtgr_frame** frame);
int rtgr_render_frame(rtgr_frame* frame, const rtgr_params* params,
                      const rtgr_object* objs, int n_objs, const rtgr_camera* cam, rtgr_stats* stats);
/* trace_rays (src:483) on ONE Pixel canvas shared by all participants of the frame: rays are read from and rgb is
 * written into `pixels` in place (the contract of rtgr_trace_canvas), the rays being drawn from the frame's shared
 * queue.  `pixels` (ni x nj x 88 B, the frame's ni/nj) must be the SAME physical page-locked host array in every
 * participant: other devices of this context see it as is; other processes map the same POSIX shared memory and
 * page-lock their mapping with rtgr_host_register.  The result is one assembled host canvas with no gather step.
 * Same protocol as rtgr_render_frame (one call per participant and frame, the caller's barrier between frames). */
int rtgr_trace_canvas_frame(rtgr_frame* frame, const rtgr_params* params, const rtgr_object* objs, int n_objs,
                            rtgr_pixel* pixels, int ni, int nj, rtgr_stats* stats);
/* Optional hint: how many GPUs (all processes together) work on this frame.  Lets the library pick the way results
 * are written that suits the crowd (rtgr_trace_canvas_frame: from four GPUs on, the colours of a patch go into the
 * host canvas as whole rows, because that many GPUs' 24-byte writes are more than a host memory system absorbs).
 * Results never depend on it. */
int rtgr_frame_set_participants(rtgr_frame* frame, int n);
/* Copy the image (nj x ni x 3, PNG order as rtgr_render's rgb8) to the host / zero it. */
int rtgr_frame_read(rtgr_frame* frame, uint8_t* rgb8);
int rtgr_frame_clear(rtgr_frame* frame);
void rtgr_frame_close(rtgr_frame* frame);

/* Device-side make_canvas alone (src:458-478): fills pos and normal of ni*nj pixels,
 * rgb = 0. */
int rtgr_make_canvas(rtgr_ctx* ctx, const rtgr_params* params, const rtgr_camera* cam,
                     rtgr_pixel* pixels);

/* Ray paths: as rtgr_trace_pixels for n rays given by their initial states (n x 8: position and null
 * 4-velocity), but every accepted step of the integrator is returned -- what the reference's solve
 * stores with its default save_everystep although trace_rays (src:502-505) only
 * reads sol[end].  Point k of ray i is the 9 doubles (lambda, x^0..3, u^0..3) at
 * paths[(i*max_points + k)*9]: k = 0 is the initial state, the last point is the state the ray was
 * coloured at (the interpolated event state, src:504).  npoints[i] is the number of points the ray
 * has; if it exceeds max_points the first max_points-1 and the last one are kept.  Unused slots are
 * zero.  Runs on the first device of the context. */
int rtgr_trace_paths(rtgr_ctx* ctx, const rtgr_params* params,
                     const rtgr_object* objs, int n_objs,
                     const double* states0, int64_t n, int32_t max_points,
                     double* paths, int32_t* npoints,
                     double* final_state, int32_t* obj_id, int32_t* status, rtgr_stats* stats);

/* ---- user-supplied metrics ------------------------------------------------------------- */

/* The reference accepts ANY callable metric(x) -> 4x4 matrix (trace_rays(metric, ...), src:483;
 * dmetric/christoffel differentiate through it with the Dual type, src:298-331).  Here the metric
 * is CUDA C++ source compiled at run time (NVRTC, sm_100a) together with the library's own
 * integrator/event/colouring code:
 *
 *     template <class T>
 *     __device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par) { ... }
 *
 * T is double (make_canvas) or a dual number carrying d/dx^0..3 (geodesic right-hand side); the
 * operator set of the reference's Dual type is available (+ - * / sqrt pow2 pow3 pow4 powi abs sin
 * cos exp log atan atan2 acos asin cbrt; see csrc/rtgr_generic.cuh).  The returned id (>=
 * RTGR_USER_METRIC_BASE) is used as rtgr_params.metric in every other entry point; M, a and
 * r_formula are ignored for it, `par` are up to 16 doubles set by rtgr_metric_set_params.
 * On a compile error the diagnostics are in rtgr_last_error().  rtgr_metric_check compiles only (no
 * device needed) and copies the compiler log into `log`.
 *
 * Two optional declarations in the source make the right-hand side cheaper (both are promises of the
 * author; nothing checks them):
 *   - a line `#pragma rtgr stationary`: g does not depend on x[0].  The duals then carry d/dx^1..3 only.
 *   - a metric of Kerr-Schild form  g_ab = eta_ab + f k_a k_b  (eta = diag(-1,1,1,1); any scalar f and any
 *     covector k, null or not) may define, INSTEAD of rtgr_user_metric,
 *
 *         template <class T>
 *         __device__ void rtgr_user_kerr_schild(const T x[4], T& f, T k[4], const double* par) { ... }
 *
 *     The geodesic acceleration is then evaluated in closed form from f, k and their derivatives
 *     (Sherman-Morrison inverse; no 4x4 derivative sets, no matrix inverse): the reference's own
 *     kerr_schild written this way (metrics/kerr_schild_form.cu) runs within 2x of the built-in kernel. */
#define RTGR_USER_METRIC_BASE 16
int rtgr_metric_compile(rtgr_ctx* ctx, const char* source, int32_t* metric_id);
int rtgr_metric_set_params(rtgr_ctx* ctx, int32_t metric_id, const double* par, int n);
int rtgr_metric_release(rtgr_ctx* ctx, int32_t metric_id);
int rtgr_metric_check(const char* source, char* log, uint64_t log_capacity);

/* ---- unit-test hooks ---------------------------------------------------------------- */

/* derivs[i] = geodesic(states[i], metric, lambda) (src:367-370), n x 8 each. */
int rtgr_rhs_batch(rtgr_ctx* ctx, const rtgr_params* params,
                   const double* states, int64_t n, double* derivs);

/* ---- resident-data variants used for kernel-only timing ----------------------------- */

/* Upload a pixel buffer once and keep it in HBM; rtgr_trace_resident then runs the trace
 * on it without any host<->device traffic (results stay on the device until fetched). */
int rtgr_upload_pixels(rtgr_ctx* ctx, const rtgr_pixel* pixels, int64_t n);
int rtgr_trace_resident(rtgr_ctx* ctx, const rtgr_params* params,
                        const rtgr_object* objs, int n_objs, rtgr_stats* stats);
/* As rtgr_render_tiles but nothing is copied back to the host. */
int rtgr_render_resident(rtgr_ctx* ctx, const rtgr_params* params,
                         const rtgr_object* objs, int n_objs, const rtgr_camera* cam,
                         int tile_offset, int tile_stride, rtgr_stats* stats);

/* Register-resident DFMA-chain microbenchmark: returns the measured FP64 FMA rate of
 * device `dev_index` of the context in TFLOP/s (an FMA = 2 flops).  This is the roofline
 * denominator ("self-measured FP64 DFMA peak"). */
int rtgr_fp64_peak(rtgr_ctx* ctx, int dev_index, double* tflops, double* sm_clock_mhz);
/* The same register-resident chains with a chosen operand mix; every mode returns "DFMA-equivalent"
 * TFLOP/s = 2 x thread-instructions/s, so all modes compare directly with mode 1 (the pipe limit):
 *   1 DFMA a=a*C1+C2 (1 register operand; = rtgr_fp64_peak)   2 DFMA a=a*b+C (2 register operands)
 *   3 DFMA a=b*c+a (3 distinct register operands)              4 DMUL a=a*b    5 DADD a=a+b
 *   6 DMUL a=a*C    7 DFMA a=b*b+a    8 DFMA a=b*c+a with b shared between neighbouring instructions
 *   9 alternating mode-3 DFMA and mode-4 DMUL
 * Modes >= 2 expose the register-file operand-read limit that general FP64 code runs into.
 * Modes 10 / 11 / 12: mode-2 DFMAs with 1 / 3 independent integer instructions per DFMA, and mode-3 DFMAs with 1
 * (do instructions of other pipes take issue or register-read bandwidth from the FP64 pipe?).
 * Modes 100 + 10*log2(chains) + w (chains in 1,2,4,8; w = 1..8 warps per scheduler, one block per SM) run
 * `chains` independent dependent-DFMA chains per thread: the latency / parallelism the pipe needs
 * (1 chain, 1 warp: rate = 1/latency). */
int rtgr_fp64_microbench(rtgr_ctx* ctx, int dev_index, int mode, double* tflops, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* RAYTRACEGR_CUDA_H */
