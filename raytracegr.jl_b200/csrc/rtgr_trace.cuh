// The per-lane ray state machine: refill -> init (initial dt, A.4) -> Tsit5 step attempts with
// error control and event detection -> finalisation (root-find, classification, colouring,
// output).  One lane owns one ray at a time; a lane that finishes immediately pulls the next ray
// from the queue, so lanes of a warp run rays of different ages but always execute the same
// six-RHS pass together.  New rays use the first two RHS slots of a pass for the two evaluations
// the initial-dt heuristic needs, so joining costs the other lanes nothing.
//
// The stage loop is ROLLED (one copy of the RHS in the instruction stream) and the seven stage
// accelerations live in a per-thread column of shared memory (`Acc`), not in registers: that keeps
// the hot loop inside the instruction cache and the kernel at <= 128 registers, i.e. 4 warps per
// scheduler to cover the FP64 pipe latency.
//
// `Sched` supplies the warp-level pieces (vote, work queue) and `Acc` the stage-acceleration
// store; the CUDA ones are in raytracegr_cuda.cu, trivial single-lane ones in tests/host_shim.cpp.
#pragma once
#include "rtgr_core.cuh"

namespace rtgr {

enum JobMode { JOB_PIXELS = 0, JOB_RENDER = 1 };

// One launch worth of work and its output buffers (any output may be null).
struct Job {
    int32_t mode;
    int32_t tiles_x, tile_offset, tile_stride;  // render mode: tile selection
    int32_t queue_scope;                        // 0: the queue head is this GPU's own (device-scope atomics);
                                                // 1: it is shared with other GPUs / processes (rtgr_frame: system scope)
    int64_t total;                              // number of queue ordinals to hand out
    const int32_t* tile_order;                  // render mode: permutation of the tile ids (or null)
    const double* pixels_in;                    // n x 11 AoS (pos, normal, rgb): rays are read from it when set
                                                // (always in pixels mode; render mode: instead of make_canvas)
    uint8_t* rgb8;                              // render mode: nj x ni x 3
    int32_t stage_rgb8;                         // launch the patch-staging kernel (host side: stage_rgb8_wanted)
    int32_t stage_canvas;                       // canvas mode with rgb_f64 = pixels_in + 8, stride 11: whole patches go back
    double* rgb_f64;                            // n x rgb_stride (3 = compact, 11 = the rgb field of a Pixel array)
    int32_t rgb_stride;
    double* final_state;                        // n x 8
    int32_t* obj_id;
    int32_t* status;
    int32_t* nsteps;
    // ray paths (rtgr_trace_paths): point k of ray `pix` = (lambda, x^0..3, u^0..3) at paths[(pix*max_points + k)*9]
    double* paths;
    int32_t* npoints;
    int32_t max_points;
};

// One point of a ray's path.  Points beyond max_points - 1 are dropped, except the last one of the ray
// (is_last), which always lands in the final slot.
RTGR_HD void record_point(const Job& job, int64_t pix, int k, bool is_last, double lam, const double* xs, const double* us) {
    if (k >= job.max_points - 1) { if (!is_last) return; k = job.max_points - 1; }
    double* o = job.paths + (pix * int64_t(job.max_points) + k) * 9;
    o[0] = lam;
    for (int c = 0; c < 4; ++c) { o[1 + c] = xs[c]; o[5 + c] = us[c]; }
}

// per-lane work counters (32 bit is ample for one lane's share of a launch; widened when reduced)
struct Counters {
    unsigned rays, attempts, accepted, rejected;
#ifdef RTGR_PASS_STATS   /* developer statistics: how many passes of a warp ran the start-up / finalisation code */
    unsigned passes = 0, init_passes = 0, fin_passes = 0;
#endif
};

enum LaneMode { L_IDLE = 0, L_FIN = 1, L_STEP = 2, L_STEP_NEAR = 3, L_DONE = 4 };

template <int METRIC, int RFORM>
RTGR_HD void accel(const SceneConst& sc, const double y[8], double A[4]) {
    if (METRIC == RTGR_MINKOWSKI) {
        A[0] = A[1] = A[2] = A[3] = 0.0;
#ifdef RTGR_USER_METRIC
    } else if (METRIC == METRIC_USER) {
        rtgr_ad::user_accel(sc.user_par, y, A);
#endif
    } else if (RFORM >= RFORM_A0) {
        ks_accel_a0<RFORM - RFORM_A0>(sc, y[1], y[2], y[3], y[4], y[5], y[6], y[7], A);
    } else {
        ks_accel<RFORM>(sc, y[1], y[2], y[3], y[4], y[5], y[6], y[7], A);
    }
}

// |dt| > eps (= dtmin) and finite, decided on the high word alone (a dt within 2^-20 relative of eps itself or a
// denormal counts as too small: both end the ray with DT_MIN, as a step of that size would a moment later)
RTGR_HD bool dt_in_range(double dt) {
    return (abs_hi_word(dt) - 0x3cb00001u) < (0x7ff00000u - 0x3cb00001u);
}

// Queue ordinal -> pixel index (or -1 for the empty part of a border tile).
RTGR_HD int64_t ordinal_to_pixel(const SceneConst& sc, const Job& job, int64_t ord, int& pi, int& pj) {
    if (job.mode == JOB_PIXELS) { pi = 0; pj = 0; return ord; }
    const int64_t m = ord >> 10;
    const int w = int(ord & 1023);
    int64_t t = job.tile_offset + m * job.tile_stride;
    if (job.tile_order) t = job.tile_order[t];   // expensive tiles first (see tile_order_by_impact)
    const int ty = int(t / job.tiles_x), tx = int(t % job.tiles_x);
    const int sub = w >> 5, l = w & 31;
    pi = tx * RTGR_TILE_W + (sub & 3) * 8 + (l & 7);
    pj = ty * RTGR_TILE_H + (sub >> 2) * 4 + (l >> 3);
    if (pi >= sc.ni || pj >= sc.nj) return -1;
    return int64_t(pi) + int64_t(pj) * sc.ni;
}

// Root of theta -> min_distance(x(theta)) inside [lo, hi] (A.5).  `lo` always keeps the sign the
// ray had at the start of the step and the bracket is driven to collapse (Illinois-modified regula
// falsi with a bisection safeguard), so the value returned is the last point before the crossing.
template <class F>
RTGR_HD double event_root(F&& cond_at, double lo, double hi, double sgn0) {
    double clo = cond_at(lo), chi = cond_at(hi);
    if (chi == 0.0) return hi;
    int side = 0;
    for (int it = 0; it < 100; ++it) {
        if (!(hi - lo > 4.440892098500626e-16 * hi)) break;
        double mid = lo - clo * (hi - lo) * fast_rcp(chi - clo);
        if (!(mid > lo && mid < hi) || (it % 3) == 2) mid = lo + 0.5 * (hi - lo);
        if (!(mid > lo && mid < hi)) break;
        const double cm = cond_at(mid);
        if (cm == 0.0) return mid;
        if (sgn0 * cm > 0.0) {
            lo = mid; clo = cm;
            if (side == -1) chi *= 0.5;
            side = -1;
        } else {
            hi = mid; chi = cm;
            if (side == +1) clo *= 0.5;
            side = +1;
        }
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// Rarely executed pieces, kept OUT OF LINE so that the per-step loop stays small enough for the
// instruction caches (L1.5 is 32 KB; fetch stalls were the top stall reason with everything
// inlined).  Arguments travel by value in registers.
// ---------------------------------------------------------------------------------------------
struct InitHead { double dt0, d1; };

// initial dt, first half (A.4): d0, d1, dt0 from u0 = (x,u) and f0 = (u, A0)
RTGR_HD InitHead init_dt_head(const SceneConst& sc, Vec4 x, Vec4 u, Vec4 A0) {
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < 4; ++c) {
        const double iskx = fast_rcp(fma(fabs(x.v[c]), sc.reltol, sc.abstol));
        const double isku = fast_rcp(fma(fabs(u.v[c]), sc.reltol, sc.abstol));
        const double ax = x.v[c] * iskx, au = u.v[c] * isku;
        const double bx = u.v[c] * iskx, bu = A0.v[c] * isku;
        s0 = fma(ax, ax, s0); s0 = fma(au, au, s0);
        s1 = fma(bx, bx, s1); s1 = fma(bu, bu, s1);
    }
    // d0 = sqrt(s0/8), d1 = sqrt(s1/8)
    double d0, d1;
    fast_rsqrt(s0 * 0.125, &d0);
    const double id1 = fast_rsqrt(s1 * 0.125, &d1);
    InitHead o;
    o.d1 = d1;
    o.dt0 = (d0 < 1e-5 || !(d1 >= 1e-5)) ? 1e-6 : (d0 * id1) * 0.01;
    o.dt0 = fmin(o.dt0, sc.dtmax);
    return o;
}

// second half: d2 from f1 - f0 = (du, dA); returns the initial dt
RTGR_HD double init_dt_tail(const SceneConst& sc, Vec4 x, Vec4 u, Vec4 du, Vec4 dA, double dt0, double d1) {
    double s2 = 0.0;
    for (int c = 0; c < 4; ++c) {
        const double iskx = fast_rcp(fma(fabs(x.v[c]), sc.reltol, sc.abstol));
        const double isku = fast_rcp(fma(fabs(u.v[c]), sc.reltol, sc.abstol));
        const double ex = du.v[c] * iskx, eu = dA.v[c] * isku;
        s2 = fma(ex, ex, s2); s2 = fma(eu, eu, s2);
    }
    double d2 = 0.0;
    if (s2 > 0.0) { fast_rsqrt(s2 * 0.125, &d2); d2 *= fast_rcp(dt0); }
    const double dm = fmax(d1, d2);
    // 10^(-(2 + log10 dm)/5)
    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : exp(-0.2 * (4.605170185988092 + log(dm)));
    return fmin(fmin(100.0 * dt0, dt1), sc.dtmax);
}

// Minkowski initial dt, reference operation order (f1 == f0, so d2 == 0)
RTGR_HD double init_dt_flat(const SceneConst& sc, Vec4 x, Vec4 u) {
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < 4; ++c) {   // state order: positions, then velocities
        const double skx = RTGR_ADD(sc.abstol, RTGR_MUL(fabs(x.v[c]), sc.reltol));
        const double ax = x.v[c] / skx, bx = u.v[c] / skx;
        s0 = RTGR_ADD(s0, RTGR_MUL(ax, ax));
        s1 = RTGR_ADD(s1, RTGR_MUL(bx, bx));
    }
    for (int c = 0; c < 4; ++c) {
        const double sku = RTGR_ADD(sc.abstol, RTGR_MUL(fabs(u.v[c]), sc.reltol));
        const double au = u.v[c] / sku;
        s0 = RTGR_ADD(s0, RTGR_MUL(au, au));   // f0's velocity part is zero
    }
    const double d0 = sqrt(s0 / 8.0);
    const double d1 = sqrt(s1 / 8.0);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
    dt0 = fmin(dt0, sc.dtmax);
    const double dm = d1;
    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
    return fmin(fmin(100.0 * dt0, dt1), sc.dtmax);
}

template <int METRIC, int RFORM>
RTGR_NOINLINE Vec8 canvas_pixel_ool(const SceneConst& sc, int pi, int pj) {
    Vec8 o;
    canvas_pixel<METRIC, RFORM>(sc, pi, pj, o.v, o.v + 4);
    return o;
}

// position on the dense output at theta
template <int METRIC>
RTGR_HD void dense_pos(const double x[4], const double u[4], double dt, const double p[4][4], double th, double q[4]) {
    if (METRIC != RTGR_MINKOWSKI) {
        for (int c = 0; c < 4; ++c) q[c] = poly_eval(x[c], p[c], th);
    } else {
        double b[7];
        dense_weights(th, b);
        for (int c = 0; c < 4; ++c) q[c] = flat_dense_x(x[c], u[c], dt, b);
    }
}

struct ScanOut { int event; double lo, hi; };

// Sample the dense output at the interior points theta_i = i/(np-1) (A.5); first sign change wins.
template <int METRIC, class Acc>
RTGR_NOINLINE ScanOut interior_scan(const SceneConst& sc, Acc acc, Vec4 x, Vec4 u, double dt, double s0) {
    double p[4][4];
    if (METRIC != RTGR_MINKOWSKI) dense_x_poly(u.v, acc, dt, p);
    ScanOut o;
    o.event = 0; o.lo = 0.0; o.hi = 1.0;
    double prev = 0.0;
#pragma unroll 1
    for (int i = 1; i <= sc.interp_points - 2; ++i) {
        const double th = sc.theta[i];
        double q[4];
        dense_pos<METRIC>(x.v, u.v, dt, p, th, q);
        const double ci = min_distance_q(sc, q[0], q[1], q[2], q[3]);
        if (s0 * ci < 0.0) { o.event = 1; o.lo = prev; o.hi = th; break; }
        prev = th;
    }
    return o;
}

// The RGB8 image (PNG order: row = j, column = i, i.e. the canvas's own linear order i + j*ni): one pixel, stored
// byte by byte.  What a scheduler without patch staging does, and what the staging one falls back to.
RTGR_HD void store_rgb8_direct(const Job& job, int64_t pix, uint32_t rgb) {
    uint8_t* o = job.rgb8 + 3 * pix;
    o[0] = uint8_t(rgb); o[1] = uint8_t(rgb >> 8); o[2] = uint8_t(rgb >> 16);
}

// Event root-find (if any), classification, colouring and the output stores of one finished ray.  A staging
// scheduler (Sched::STAGE) takes the RGB8 pixel: it collects a warp's 8x4-pixel patch in shared memory and writes
// it out in 8-byte stores; the return value is then the staging slot this ray COMPLETED (0 or 1), else -1.
template <int METRIC, class Acc, class Sched>
RTGR_NOINLINE int finalize_ray(const SceneConst& sc, const Job& job, Acc acc, Vec4 x, Vec4 u, Vec8 y, double dt,
                                double th_lo, double th_hi, double cprev, double c_new, int have_root,
                                int64_t pix, int status, int nacc, double tstep) {
    double fs[8];
    double th_fin = 0.0;     // the ray ends at lambda = tstep + th_fin * dt
    for (int c = 0; c < 4; ++c) { fs[c] = x.v[c]; fs[4 + c] = u.v[c]; }
    if (have_root) {
        const double sgn0 = (cprev > 0.0) ? 1.0 : -1.0;
        double p[4][4];
        if (METRIC != RTGR_MINKOWSKI) dense_x_poly(u.v, acc, dt, p);
        // Which objects change sign across the bracket?  Almost always exactly one: then the
        // condition min_o d_o has the same root as that single d_o inside the bracket as long as the
        // others stay on their side, which they do when they do not change sign.
        double qlo[4], qhi[4];
        dense_pos<METRIC>(x.v, u.v, dt, p, th_lo, qlo);
        dense_pos<METRIC>(x.v, u.v, dt, p, th_hi, qhi);
        if (th_lo == 0.0) for (int c = 0; c < 4; ++c) qlo[c] = x.v[c];
        if (th_hi == 1.0) for (int c = 0; c < 4; ++c) qhi[c] = y.v[c];
        int ncross = 0, ocross = 0;
        {
            const double nlo = fma(qlo[1], qlo[1], fma(qlo[2], qlo[2], qlo[3] * qlo[3]));
            const double nhi = fma(qhi[1], qhi[1], fma(qhi[2], qhi[2], qhi[3] * qhi[3]));
#pragma unroll 1
            for (int o = 0; o < sc.n_objs; ++o) {
                const double a = obj_distance_q(sc, o, nlo, qlo[0], qlo[1], qlo[2], qlo[3]);
                const double b = obj_distance_q(sc, o, nhi, qhi[0], qhi[1], qhi[2], qhi[3]);
                if ((a > 0.0) != (b > 0.0) || a == 0.0 || b == 0.0) { ++ncross; ocross = o; }
            }
        }
        const bool single = (ncross == 1) && (cprev > 0.0);
        auto cond_at = [&](double th) -> double {
            if (th == 1.0) return c_new;
            if (th == 0.0) return cprev;
            double q[4];
            dense_pos<METRIC>(x.v, u.v, dt, p, th, q);
            if (single) {
                const double n2 = fma(q[1], q[1], fma(q[2], q[2], q[3] * q[3]));
                return obj_distance_q(sc, ocross, n2, q[0], q[1], q[2], q[3]);
            }
            return min_distance_q(sc, q[0], q[1], q[2], q[3]);
        };
        const double th_star = event_root(cond_at, th_lo, th_hi, sgn0);
        th_fin = th_star;
        if (th_star == 1.0) {
            for (int c = 0; c < 8; ++c) fs[c] = y.v[c];
        } else if (th_star > 0.0) {
            dense_pos<METRIC>(x.v, u.v, dt, p, th_star, fs);
            if (METRIC != RTGR_MINKOWSKI) dense_u(u.v, acc, dt, th_star, fs + 4);
        }
    }
    double col[3];
    const int omin = classify_color(sc, fs, col);
    int completed = -1;
    if (job.rgb_f64) {
        if (Sched::PREFETCH && job.stage_canvas) completed = Sched::put_rgbf(sc, job, int32_t(pix), col);
        else { for (int c = 0; c < 3; ++c) job.rgb_f64[int64_t(job.rgb_stride) * pix + c] = col[c]; }
    }
    if (job.rgb8) {
        const uint32_t rgb = uint32_t(quantize8(col[0])) | (uint32_t(quantize8(col[1])) << 8) | (uint32_t(quantize8(col[2])) << 16);
        if (Sched::STAGE) completed = Sched::put_rgb8(sc, job, int32_t(pix), rgb);
        else store_rgb8_direct(job, pix, rgb);
    }
    if (job.final_state) { for (int c = 0; c < 8; ++c) job.final_state[8 * pix + c] = fs[c]; }
    if (job.obj_id) job.obj_id[pix] = omin;
    if (job.status) job.status[pix] = status;
    if (job.nsteps) job.nsteps[pix] = nacc;
    if (job.paths) {
        // an event ends the ray at the interpolated state (its own point); otherwise the last accepted step
        // (already recorded) is the end
        if (have_root) record_point(job, pix, nacc, true, fma(th_fin, dt, tstep), fs, fs + 4);
        // a ray that ends without an event (lambda1, maxiters, dtmin, NaN) after more accepted steps than the buffer
        // holds: its last state goes into the final slot too (the truncation contract of rtgr_trace_paths)
        else if (nacc >= job.max_points - 1) record_point(job, pix, nacc, true, tstep, fs, fs + 4);
        if (job.npoints) job.npoints[pix] = nacc + 1;
    }
    return completed;
}

RTGR_HD Vec4 mk4(const double* a) { Vec4 r; for (int c = 0; c < 4; ++c) r.v[c] = a[c]; return r; }
RTGR_HD Vec8 mk8(const double* a) { Vec8 r; for (int c = 0; c < 8; ++c) r.v[c] = a[c]; return r; }

// Start of a ray (rare: once per ray, out of line): the two right-hand sides of the initial-dt heuristic (A.4),
// the first stage acceleration (FSAL slot 0) and the distance at the start point.
struct InitOut { double dt, cprev; int status; };   // status >= 0: the ray ends before its first step

template <int METRIC, int RFORM, class AccB>
RTGR_NOINLINE InitOut init_ray(const SceneConst& sc, AccB acc, Vec4 x, Vec4 u) {
    InitOut o;
    o.status = -1;
    if (METRIC == RTGR_MINKOWSKI) {
        o.dt = init_dt_flat(sc, x, u);
    } else {
        double y[8], A0[4], A1[4];
        for (int c = 0; c < 4; ++c) { y[c] = x.v[c]; y[4 + c] = u.v[c]; }
        accel<METRIC, RFORM>(sc, y, A0);                       // f(u0): also k1 of the first step
        acc.store(0, A0);
        const InitHead ih = init_dt_head(sc, x, u, mk4(A0));
        for (int c = 0; c < 4; ++c) { y[c] = fma(ih.dt0, u.v[c], x.v[c]); y[4 + c] = fma(ih.dt0, A0[c], u.v[c]); }
        accel<METRIC, RFORM>(sc, y, A1);                       // f(u0 + dt0 f0)
        Vec4 du, dA;
        for (int c = 0; c < 4; ++c) { du.v[c] = y[4 + c] - u.v[c]; dA.v[c] = A1[c] - A0[c]; }
        o.dt = init_dt_tail(sc, x, u, du, dA, ih.dt0, ih.d1);
    }
    bool bad = false;
    for (int c = 0; c < 4; ++c) bad = bad || !(x.v[c] == x.v[c]) || !(u.v[c] == u.v[c]);
    o.cprev = min_distance_q(sc, x.v[0], x.v[1], x.v[2], x.v[3]);
    if (bad) o.status = RTGR_STATUS_NONFINITE;
    else if (!(sc.lambda0 < sc.lambda1)) o.status = RTGR_STATUS_LAMBDA_END;
    return o;
}

// t and dt so small against lambda1 that the coming step ends well below it: t < lambda1/2 and dt < lambda1/4 on
// the high words (t1_half_hi = INT32_MIN switches the shortcut off, e.g. for a span that is not positive)
RTGR_HD bool step_far_from_end(const SceneConst& sc, double t, double dt) {
    return hi_word_signed(t) < sc.t1_half_hi && hi_word_signed(dt) < sc.t1_quarter_hi;   // (lambda0 >= 0, so t >= 0)
}

// signed order of the high words: a > b whenever this holds (a sufficient test; used with b >= 0)
RTGR_HD bool hi_gt(double a, double b) { return int32_t(hi_word(a)) > int32_t(hi_word(b)); }

template <int METRIC, int RFORM, class Sched, class Acc, bool PATHS = false>
RTGR_HD void trace_loop(const SceneConst& sc, const StageTab& T, const Job& job, Sched& sched, Acc& acc,
                        Counters& cnt) {
    constexpr bool FLAT = (METRIC == RTGR_MINKOWSKI);
    // The PI controller's step-size FACTOR (A.3): FP64 log/exp for Minkowski (whose step sequence is reproduced
    // step for step) and, in a -DRTGR_CONTROLLER_FP64 build, for every metric; otherwise lg2/ex2.approx on the
    // FP32/SFU pipes (controller_inv_q_fast).  Accept/reject and all state arithmetic are FP64 either way.
#ifdef RTGR_CONTROLLER_FP64
    constexpr bool CTL64 = true;
#else
    constexpr bool CTL64 = FLAT;
#endif
    // the time coordinate of the intermediate stages 2..6 is only formed when the right-hand side can read it
    constexpr bool STAGE_T = (METRIC == METRIC_USER);
    // ---- lane state.  The hot loop is bound by instruction dispatch and the kernel by its 128 registers, so the
    // ---- state that lives across a pass is kept small: what the common step does not read is not carried.
    double x[4], u[4];   // state at the start of the current step
    double y[8];         // stage state / candidate new state
    double dt = 0.0, t = 0.0, lqold = LOG_QOLDINIT;
    float lqold2 = LOG2_QOLDINIT_F;      // Kerr-Schild path: log2(qold) (see controller_inv_q_fast)
    // min_distance at the start of the step is carried as its (signed) high word only: that is all the per-step
    // fast path reads; the rare steps that need the value recompute it from x (bit-identical: same function, same point)
    int32_t cprev_hi = 0;
    int32_t pix = -1;                    // (the entry points reject frames of 2^31 rays or more)
    int mode = L_IDLE;                   // L_STEP_NEAR = L_STEP with the lambda1 rules armed (see step_far_from_end)
    int left = 0;                        // step attempts left before maxiters
    unsigned nrej = 0;                   // attempts that were not accepted (rare; read only when a ray ends)
#pragma unroll
    for (int c = 0; c < 4; ++c) { x[c] = 0.0; u[c] = 0.0; }
#pragma unroll
    for (int c = 0; c < 8; ++c) y[c] = 1.0;
    {
        const double z4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < 7; ++i) acc.store(i, z4);
    }
    const double t1 = sc.lambda1;

    for (;;) {
        // how a ray that ends in this pass ends: defined on the (rare) paths that set mode = L_FIN, read by the
        // finalisation at the bottom -- deliberately NOT initialised here, so that nothing is carried through the pass
        int fin_status;
        int have_root;
        double th_lo, th_hi, c0, c1;
        // ============ refill + start of the new rays (rare: a few per cent of the passes) ============
        if (sched.any(mode == L_IDLE)) {
            const bool idle = (mode == L_IDLE);
            const int64_t ord = sched.fetch(idle, job);
            // RGB8 patch staging: write out the patches completed by the rays that just ended (their lanes carry the
            // slot in `pix`) and open a slot for a newly drawn chunk
            if (Sched::STAGE) sched.stage_refill(job, idle ? pix : -1);
            // rays from a Pixel array: the warp reads a chunk's 32 rays together when it draws the chunk
            double ray[8];
            if (Sched::PREFETCH) sched.take_rays(job, idle, ord, idle ? pix : -1, ray);
            if (idle) {
                if (ord >= job.total) {
                    mode = L_DONE;
                } else {
                    int pi, pj;
                    pix = int32_t(ordinal_to_pixel(sc, job, ord, pi, pj));
                    if (pix >= 0) {
                        if (Sched::PREFETCH) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) { x[c] = ray[c]; u[c] = ray[4 + c]; }   // src:492-496
                        } else if (job.pixels_in) {
                            const double* px = job.pixels_in + 11 * int64_t(pix);
#pragma unroll
                            for (int c = 0; c < 4; ++c) { x[c] = px[c]; u[c] = px[4 + c]; }   // src:492-496
                        } else {
                            const Vec8 xu = canvas_pixel_ool<METRIC, RFORM>(sc, pi, pj);
#pragma unroll
                            for (int c = 0; c < 4; ++c) { x[c] = xu.v[c]; u[c] = xu.v[4 + c]; }
                        }
                        const InitOut io = init_ray<METRIC, RFORM, typename Acc::Backing>(sc, acc.backing(), mk4(x), mk4(u));
                        dt = io.dt; cprev_hi = hi_word_signed(io.cprev); t = sc.lambda0;
                        lqold = LOG_QOLDINIT; lqold2 = LOG2_QOLDINIT_F; left = sc.maxiters; nrej = 0;
                        if (PATHS) record_point(job, pix, 0, false, t, x, u);
                        mode = step_far_from_end(sc, t, dt) ? L_STEP : L_STEP_NEAR;
                        have_root = 0; th_lo = 0.0; th_hi = 1.0; c0 = 0.0; c1 = 0.0;
                        if (io.status >= 0) { mode = L_FIN; fin_status = io.status; }
                        else if (left <= 0) { mode = L_FIN; fin_status = RTGR_STATUS_MAXITERS; }
                        else if (!dt_in_range(dt)) { mode = L_FIN; fin_status = (dt == dt) ? RTGR_STATUS_DT_MIN : RTGR_STATUS_NONFINITE; }
                        cnt.rays += 1;
                    }
                }
            }
            if (sched.all(mode == L_DONE)) break;
        }

        // =============================== pre-step ===============================
        // "Never step past lambda1": only when t or dt are a sizeable fraction of lambda1 -- in practice never, the
        // rays end on an object long before -- can the clamp bind; the exact rule runs there.  (The other pre-step
        // checks of the solver, maxiters and dt >= dtmin, are made right after the values changed: at the end of
        // the previous pass.)
        if (mode == L_STEP_NEAR) dt = min_mixed(dt, t1 - t);   // (both positive here)
        const bool stepping = (mode == L_STEP) || (mode == L_STEP_NEAR);
#ifdef RTGR_PASS_STATS
        cnt.passes += 1;
#endif

        double msq = 0.0;
        uint32_t amax_hi = 0;
        const double dt2 = dt * dt;
        if (!FLAT) {
            // ---- six right-hand sides: stages 2..7 ----
            // UNROLLED: six copies of the right-hand side (the hot loop is ~35 KB of code; measured on B200 the
            // instruction fetch keeps up, and the kernel is bound by instruction dispatch / register reads, where
            // the switch, the loop counters and the per-stage address arithmetic of a rolled loop cost 5 % of the frame)
#ifdef RTGR_ROLLED_STAGES
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int s = 2; s <= 7; ++s) {
                switch (s) {
                    case 2: stage_state<2, STAGE_T>(T, x, u, acc, dt, dt2, y); break;
                    case 3: stage_state<3, STAGE_T>(T, x, u, acc, dt, dt2, y); break;
                    case 4: stage_state<4, STAGE_T>(T, x, u, acc, dt, dt2, y); break;
                    case 5: stage_state<5, STAGE_T>(T, x, u, acc, dt, dt2, y); break;
                    case 6: stage_state<6, STAGE_T>(T, x, u, acc, dt, dt2, y); break;
                    default: stage_state<7, true>(T, x, u, acc, dt, dt2, y); break;
                }
                double An[4];
                accel<METRIC, RFORM>(sc, y, An);
                acc.store(s - 1, An);
            }
            // y now holds the candidate new state (stage 7), acc[6] its acceleration
            msq = error_msq(sc, T, x, u, acc, dt, dt2, y, amax_hi);
        } else {
            // Minkowski: RHS == (u, 0) at every stage
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                y[c] = RTGR_ADD(x[c], RTGR_MUL(dt, flat_sum_a7(u[c])));
                y[4 + c] = u[c];
            }
            msq = flat_error_msq(sc, x, u, dt, y);
        }

        // ---- error control (A.2, A.3) and event detection (A.5) for the lanes that stepped ----
        {
            double lE = 0.0;
            float lE2 = 0.0f;
            const double inv_q = CTL64 ? controller_inv_q(T, msq, lqold, lE) : controller_inv_q_fast(msq, lqold2, lE2);
            // end-point distance + conservative "nothing in reach" bound along the chord (see coarse_need)
            const double umax = from_hi_word(float_bits(fmaxf(fmaxf(fabsf(hi_word_as_float(u[0])), fabsf(hi_word_as_float(u[1]))),
                                                              fmaxf(fabsf(hi_word_as_float(u[2])), fabsf(hi_word_as_float(u[3]))))) + 1u);
            const double amax = FLAT ? 0.0 : from_hi_word(amax_hi + 1u);
            const double dev = dt * fma(T.chord_dev * dt, amax, 1e-13 * umax);
            int32_t c1_hi = min_distance_q4_hi(sc, y[0], y[1], y[2], y[3]);
            const double need = coarse_need(sc, x, y, dev);
            const int32_t need_hi = hi_word_signed(need);
            // The step that needs nothing else: accepted, outside every object at both ends (all distances > need >= 0:
            // no sign change) and no object within reach in between (coarse bound: the interior samples of A.5
            // cannot change sign).  Decided on the high words: a tie goes to the exact logic below.
            const bool plain = stepping && le_one_nonneg(msq) && cprev_hi > need_hi && c1_hi > need_hi;
            bool advance = plain;
            if (stepping && !plain) {
                // ---- everything else (rare): the rules in full, on the exact values ----
                c0 = min_distance_q4(sc, x[0], x[1], x[2], x[3]);
                c1 = min_distance_q4(sc, y[0], y[1], y[2], y[3]);
                c1_hi = hi_word_signed(c1);
                th_lo = 0.0; th_hi = 1.0; have_root = 0;
                const bool accept = le_one_nonneg(msq);
                const bool ppos = is_pos(c0), pneg = is_neg(c0);
                const bool crossing = ppos ? !is_pos(c1) : (pneg ? !is_neg(c1) : false);
                // Interior samples are needed when the end points agree in sign but a visit in between cannot be
                // ruled out (or the ray started inside an object).  Two-level test: the coarse bound from the
                // minima, then the per-object chord test.
                bool need_scan = accept && !crossing && (ppos || pneg) && (sc.interp_points > 2) &&
                                 !(ppos && gt_nonneg(c0, need) && gt_nonneg(c1, need));
                if (need_scan && ppos) {
                    bool clear;
                    end_distances(sc, x, y, dev, clear);
                    need_scan = !clear;
                }
                bool event = accept && crossing;
                if (need_scan) {
                    const double s0 = ppos ? 1.0 : -1.0;
                    const ScanOut so = interior_scan<METRIC, typename Acc::Backing>(sc, acc.backing(), mk4(x), mk4(u), dt, s0);
                    if (so.event) { event = true; th_lo = so.lo; th_hi = so.hi; }
                }
                if (accept) {
                    if (event) { --left; mode = L_FIN; fin_status = RTGR_STATUS_EVENT; have_root = 1; }
                    else advance = true;
                } else if (!is_nan_bits(msq)) {
                    dt *= reject_factor(CTL64 ? lE : double(lE2) * 0.6931471805599453);   // rejected: same state, smaller step
                    --left; nrej += 1;
                    cnt.rejected += 1;
                    if (left <= 0) { mode = L_FIN; fin_status = RTGR_STATUS_MAXITERS; }
                    else if (!dt_in_range(dt)) { mode = L_FIN; fin_status = (dt == dt) ? RTGR_STATUS_DT_MIN : RTGR_STATUS_NONFINITE; }
                } else {
                    // NaN error estimate (e.g. rho < a under the as-written radius): stop the ray here
                    --left; nrej += 1;
                    mode = L_FIN; fin_status = RTGR_STATUS_NONFINITE;
                }
            }
            if (advance) {
                // accepted, no event: advance; FSAL: the last stage's acceleration opens the next step
                --left;
                const double tt = t + dt;
                t = tt;
                if (CTL64) lqold = max_nonpos(lE, LOG_QOLDINIT);   // accepted: EEst <= 1, so lE <= 0
                else lqold2 = fmaxf(lE2, LOG2_QOLDINIT_F);
                dt = dt * inv_q;
                cprev_hi = c1_hi;
#pragma unroll
                for (int c = 0; c < 4; ++c) { x[c] = y[c]; u[c] = y[4 + c]; }
                if (!FLAT) {
                    double A7[4];
                    acc.load(6, A7);
                    acc.store(0, A7);
                }
                // The solver's rules before the next attempt, in full only when one of them can bind (rare): dt is
                // capped by dtmax, lambda1 ends the ray (with the solver's snap of t to lambda1), maxiters, and dt
                // must be a finite number >= dtmin (= eps).
                const bool far = step_far_from_end(sc, t, dt);   // also false for a dt at or above ~dtmax/4 or NaN
                if (mode == L_STEP_NEAR || !far || left <= 0 || !dt_in_range(dt)) {
                    dt = min_mixed(dt, sc.dtmax);
                    if (mode == L_STEP_NEAR) t = (fabs(tt - t1) < 10.0 * 2.220446049250313e-16 * fmax(tt, t1)) ? t1 : tt;
                    mode = step_far_from_end(sc, t, dt) ? L_STEP : L_STEP_NEAR;
                    have_root = 0; th_lo = 0.0; th_hi = 1.0; c0 = 0.0; c1 = 0.0;
                    if (!(t < t1)) { mode = L_FIN; fin_status = RTGR_STATUS_LAMBDA_END; }
                    else if (left <= 0) { mode = L_FIN; fin_status = RTGR_STATUS_MAXITERS; }
                    else if (!dt_in_range(dt)) { mode = L_FIN; fin_status = (dt == dt) ? RTGR_STATUS_DT_MIN : RTGR_STATUS_NONFINITE; }
                }
                if (PATHS) record_point(job, pix, sc.maxiters - left - int(nrej), false, t, x, u);
            }
        }

        // =============================== finalisation ===============================
#ifdef RTGR_PASS_STATS
        if (sched.any(mode == L_FIN)) cnt.fin_passes += 1;
#endif
        if (mode == L_FIN) {
            const int nacc = sc.maxiters - left - int(nrej);
            const int completed = finalize_ray<METRIC, typename Acc::Backing, Sched>(sc, job, acc.backing(), mk4(x), mk4(u), mk8(y), dt,
                                                                                     th_lo, th_hi, c0, c1, have_root, pix, fin_status, nacc, t);
            if (Sched::STAGE || Sched::PREFETCH) pix = -2 - completed;   // -1: nothing to flush; -2 / -3: this ray completed slot 0 / 1
            cnt.attempts += unsigned(sc.maxiters - left);
            cnt.accepted += unsigned(nacc);
            mode = L_IDLE;
        }
    }
}

}  // namespace rtgr
