// The per-lane ray state machine: refill -> init (initial dt, A.4) -> Tsit5 step attempts with
// error control and event detection -> finalisation (root-find, classification, colouring,
// output).  One lane owns one ray at a time; a lane that finishes immediately pulls the next ray
// from the queue, so lanes of a warp run rays of different ages but always execute the same
// six-RHS pass together.  New rays use the first two RHS slots of a pass for the two evaluations
// the initial-dt heuristic needs, so joining costs the other lanes nothing.
//
// `Sched` supplies the warp-level pieces (vote, work queue); the CUDA one is in
// raytracegr_cuda.cu, a trivial single-lane one lives in tests/host_shim.cpp.
#pragma once
#include "rtgr_core.cuh"

namespace rtgr {

enum JobMode { JOB_PIXELS = 0, JOB_RENDER = 1 };

// One launch worth of work and its output buffers (any output may be null).
struct Job {
    int32_t mode;
    int32_t tiles_x, tile_offset, tile_stride;  // render mode: tile selection
    int64_t total;                              // number of queue ordinals to hand out
    const double* pixels_in;                    // pixels mode: n x 11 AoS (pos, normal, rgb)
    uint8_t* rgb8;                              // render mode: nj x ni x 3
    double* rgb_f64;                            // n x 3
    double* final_state;                        // n x 8
    int32_t* obj_id;
    int32_t* status;
    int32_t* nsteps;
};

struct Counters {
    unsigned long long rays, attempts, accepted, rejected;
};

enum LaneMode { L_IDLE = 0, L_INIT = 1, L_STEP = 2, L_FIN = 3, L_DONE = 4 };

template <int METRIC, int RFORM>
RTGR_HD void accel(const SceneConst& sc, const double y[8], double A[4]) {
    if (METRIC == RTGR_MINKOWSKI) {
        A[0] = A[1] = A[2] = A[3] = 0.0;
    } else {
        ks_accel<RFORM>(sc, y[1], y[2], y[3], y[4], y[5], y[6], y[7], A);
    }
}

// Queue ordinal -> pixel index (or -1 for the empty part of a border tile).
RTGR_HD int64_t ordinal_to_pixel(const SceneConst& sc, const Job& job, int64_t ord, int& pi, int& pj) {
    if (job.mode == JOB_PIXELS) { pi = 0; pj = 0; return ord; }
    const int64_t m = ord >> 10;
    const int w = int(ord & 1023);
    const int64_t t = job.tile_offset + m * job.tile_stride;
    const int ty = int(t / job.tiles_x), tx = int(t % job.tiles_x);
    const int sub = w >> 5, l = w & 31;
    pi = tx * RTGR_TILE_W + (sub & 3) * 8 + (l & 7);
    pj = ty * RTGR_TILE_H + (sub >> 2) * 4 + (l >> 3);
    if (pi >= sc.ni || pj >= sc.nj) return -1;
    return int64_t(pi) + int64_t(pj) * sc.ni;
}

template <int METRIC, int RFORM, class Sched>
RTGR_HD void trace_loop(const SceneConst& sc, const Job& job, Sched& sched, Counters& cnt) {
    // ---- lane state (registers) ----
    double x[4], u[4];   // state at the start of the current step
    double A[7][4];      // stage accelerations (A[0] = FSAL slope)
    double y[8];         // stage state / candidate
    double dt = 0.0, t = 0.0, lqold = LOG_QOLDINIT, cprev = 0.0;
    double dt0 = 0.0, d1 = 0.0;          // init scratch
    double th_lo = 0.0, th_hi = 1.0, c_new = 0.0;
    int64_t pix = -1;
    int pi = 0, pj = 0;
    int mode = L_IDLE, status = RTGR_STATUS_EVENT, iter = 0, nacc = 0;
    bool have_root = false;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[i][c] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) { x[c] = 0.0; u[c] = 0.0; }
#pragma unroll
    for (int c = 0; c < 8; ++c) y[c] = 1.0;

    const double t1 = sc.lambda1;

    for (;;) {
        // =============================== refill ===============================
        {
            const bool idle = (mode == L_IDLE);
            if (sched.any(idle)) {
                const int64_t ord = sched.fetch(idle);
                if (idle) {
                    if (ord >= job.total) {
                        mode = L_DONE;
                    } else {
                        pix = ordinal_to_pixel(sc, job, ord, pi, pj);
                        if (pix >= 0) {
                            if (job.mode == JOB_PIXELS) {
                                const double* px = job.pixels_in + 11 * pix;
#pragma unroll
                                for (int c = 0; c < 4; ++c) { x[c] = px[c]; u[c] = px[4 + c]; }   // src:492-496
                            } else {
                                canvas_pixel<METRIC, RFORM>(sc, pi, pj, x, u);
                            }
                            mode = L_INIT;
                            cnt.rays += 1;
                        }
                    }
                }
            }
            if (sched.all(mode == L_DONE)) break;
        }

        // =============================== pre-step ===============================
        if (mode == L_STEP) {
            dt = fmin(dt, t1 - t);                      // never step past lambda1
            ++iter;
            if (iter > sc.maxiters) { mode = L_FIN; status = RTGR_STATUS_MAXITERS; have_root = false; }
            else if (!(fabs(dt) > 2.220446049250313e-16)) {
                mode = L_FIN; have_root = false;
                status = (dt == dt) ? RTGR_STATUS_DT_MIN : RTGR_STATUS_NONFINITE;
            }
        }
        const bool stepping = (mode == L_STEP);
        const bool initing = (mode == L_INIT);
        if (stepping) cnt.attempts += 1;

        double msq = 0.0;
        if (METRIC != RTGR_MINKOWSKI) {
            double An[4];
            // ---- RHS slot 1: stage 2, or f(u0) for a new ray ----
            stage_state<2>(x, u, A, dt, y);
            if (initing) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { y[c] = x[c]; y[4 + c] = u[c]; }
            }
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) { A[1][c] = An[c]; if (initing) A[0][c] = An[c]; }
            if (sched.any(initing)) {
                if (initing) {
                    // initial dt, first half (A.4): d0, d1, dt0 and the Euler probe state
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double skx = fma(fabs(x[c]), sc.reltol, sc.abstol);
                        const double sku = fma(fabs(u[c]), sc.reltol, sc.abstol);
                        const double ax = x[c] / skx, au = u[c] / sku;
                        const double bx = u[c] / skx, bu = A[0][c] / sku;     // f0 = (u, A0)
                        s0 = fma(ax, ax, s0); s0 = fma(au, au, s0);
                        s1 = fma(bx, bx, s1); s1 = fma(bu, bu, s1);
                    }
                    const double d0 = sqrt(s0 * 0.125);
                    d1 = sqrt(s1 * 0.125);
                    dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
                    dt0 = fmin(dt0, sc.dtmax);
                }
            }
            // ---- RHS slot 2: stage 3, or f(u0 + dt0 f0) ----
            stage_state<3>(x, u, A, dt, y);
            if (initing) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { y[c] = fma(dt0, u[c], x[c]); y[4 + c] = fma(dt0, A[0][c], u[c]); }
            }
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) A[2][c] = An[c];
            if (sched.any(initing)) {
                if (initing) {
                    // second half: d2 from f1 - f0 = (dt0*A0, An - A0)
                    double s2 = 0.0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double skx = fma(fabs(x[c]), sc.reltol, sc.abstol);
                        const double sku = fma(fabs(u[c]), sc.reltol, sc.abstol);
                        const double ex = (y[4 + c] - u[c]) / skx;
                        const double eu = (An[c] - A[0][c]) / sku;
                        s2 = fma(ex, ex, s2); s2 = fma(eu, eu, s2);
                    }
                    const double d2 = sqrt(s2 * 0.125) / dt0;
                    const double dm = fmax(d1, d2);
                    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3)
                                                     : pow(10.0, -(2.0 + log10(dm)) / 5.0);
                    dt0 = fmin(fmin(100.0 * dt0, dt1), sc.dtmax);   // becomes dt at the end of the pass
                }
            }
            // ---- RHS slots 3..6: stages 4..7 ----
            stage_state<4>(x, u, A, dt, y);
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) A[3][c] = An[c];
            stage_state<5>(x, u, A, dt, y);
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) A[4][c] = An[c];
            stage_state<6>(x, u, A, dt, y);
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) A[5][c] = An[c];
            stage_state<7>(x, u, A, dt, y);      // candidate new state
            accel<METRIC, RFORM>(sc, y, An);
#pragma unroll
            for (int c = 0; c < 4; ++c) A[6][c] = An[c];
            if (stepping) msq = error_msq(sc, x, u, A, dt, y);
        } else {
            // Minkowski: RHS == (u, 0) at every stage
            if (initing) {
                double s0 = 0.0, s1 = 0.0;
                for (int c = 0; c < 4; ++c) {   // state order: positions, then velocities
                    const double skx = RTGR_ADD(sc.abstol, RTGR_MUL(fabs(x[c]), sc.reltol));
                    const double ax = x[c] / skx, bx = u[c] / skx;
                    s0 = RTGR_ADD(s0, RTGR_MUL(ax, ax));
                    s1 = RTGR_ADD(s1, RTGR_MUL(bx, bx));
                }
                for (int c = 0; c < 4; ++c) {
                    const double sku = RTGR_ADD(sc.abstol, RTGR_MUL(fabs(u[c]), sc.reltol));
                    const double au = u[c] / sku;
                    s0 = RTGR_ADD(s0, RTGR_MUL(au, au));   // f0's velocity part is zero
                }
                const double d0 = sqrt(s0 / 8.0);
                d1 = sqrt(s1 / 8.0);
                dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
                dt0 = fmin(dt0, sc.dtmax);
                const double dm = d1;   // f1 == f0, so d2 == 0
                const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3)
                                                 : pow(10.0, -(2.0 + log10(dm)) / 5.0);
                dt0 = fmin(fmin(100.0 * dt0, dt1), sc.dtmax);
            }
            if (stepping) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    y[c] = RTGR_ADD(x[c], RTGR_MUL(dt, flat_sum_a7(u[c])));
                    y[4 + c] = u[c];
                }
                msq = flat_error_msq(sc, x, u, dt, y);
            }
        }

        // =============================== end of pass ===============================
        if (initing) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) bad = bad || !(x[c] == x[c]) || !(u[c] == u[c]);
            dt = dt0; t = sc.lambda0; lqold = LOG_QOLDINIT; iter = 0; nacc = 0;
            cprev = min_distance(sc, x[0], x[1], x[2], x[3]);
            mode = L_STEP;
            if (bad) { mode = L_FIN; status = RTGR_STATUS_NONFINITE; have_root = false; }
            else if (!(t < t1)) { mode = L_FIN; status = RTGR_STATUS_LAMBDA_END; have_root = false; }
        }
        if (stepping) {
            double lE;
            const double inv_q = controller_inv_q(msq, lqold, lE);
            if (!(msq == msq)) {
                // NaN error estimate (e.g. rho < a under the as-written radius): stop the ray here
                mode = L_FIN; status = RTGR_STATUS_NONFINITE; have_root = false;
            } else if (!(msq <= 1.0)) {
                cnt.rejected += 1;
                dt *= reject_factor(lE);
            } else {
                cnt.accepted += 1; ++nacc;
                lqold = fmax(lE, LOG_QOLDINIT);
                const double dtnew = fmin(dt * inv_q, sc.dtmax);
                const double ttmp = t + dt;
                const double tnew = (fabs(ttmp - t1) < 10.0 * 2.220446049250313e-16 * fmax(ttmp, t1)) ? t1 : ttmp;
                // ---- ContinuousCallback (A.5): sign change of min_distance over the step ----
                const double c0 = cprev;
                const double c1 = min_distance(sc, y[0], y[1], y[2], y[3]);
                const double s0 = (c0 > 0.0) ? 1.0 : ((c0 < 0.0) ? -1.0 : 0.0);
                const double s1 = (c1 > 0.0) ? 1.0 : ((c1 < 0.0) ? -1.0 : 0.0);
                bool event = false;
                th_lo = 0.0; th_hi = 1.0;
                if (s0 != 0.0 && s0 * s1 <= 0.0) {
                    event = true;
                } else if (s0 != 0.0) {
                    // interior dense-output sample points theta_i = i/(np-1)
                    double p[4][4];
                    if (METRIC != RTGR_MINKOWSKI) dense_x_poly(u, A, dt, p);
                    double prev = 0.0;
                    for (int i = 1; i <= sc.interp_points - 2; ++i) {
                        const double th = sc.theta[i];
                        double q[4];
                        if (METRIC != RTGR_MINKOWSKI) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) q[c] = poly_eval(x[c], p[c], th);
                        } else {
                            double b[7];
                            dense_weights(th, b);
#pragma unroll
                            for (int c = 0; c < 4; ++c) q[c] = flat_dense_x(x[c], u[c], dt, b);
                        }
                        const double ci = min_distance(sc, q[0], q[1], q[2], q[3]);
                        if (!event && s0 * ci < 0.0) { event = true; th_lo = prev; th_hi = th; }
                        if (!event) prev = th;
                    }
                }
                if (event) {
                    mode = L_FIN; status = RTGR_STATUS_EVENT; have_root = true; c_new = c1;
                } else {
                    // accept: advance, FSAL
                    t = tnew; dt = dtnew; cprev = c1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) { x[c] = y[c]; u[c] = y[4 + c]; A[0][c] = A[6][c]; }
                    if (!(t < t1)) { mode = L_FIN; status = RTGR_STATUS_LAMBDA_END; have_root = false; }
                }
            }
        }

        // =============================== finalisation ===============================
        if (sched.any(mode == L_FIN)) {
            if (mode == L_FIN) {
                double fs[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) { fs[c] = x[c]; fs[4 + c] = u[c]; }
                if (have_root) {
                    // Root of theta -> min_distance(x(theta)) inside [th_lo, th_hi]; `lo` always keeps
                    // the sign the ray had at the start of the step, and the bracket is driven to
                    // collapse, so the state taken is the last one before the crossing (A.5).
                    const double sgn0 = (cprev > 0.0) ? 1.0 : -1.0;
                    double p[4][4];
                    if (METRIC != RTGR_MINKOWSKI) dense_x_poly(u, A, dt, p);
                    auto cond_at = [&](double th) -> double {
                        if (th == 1.0) return c_new;
                        if (th == 0.0) return cprev;
                        double q[4];
                        if (METRIC != RTGR_MINKOWSKI) {
                            for (int c = 0; c < 4; ++c) q[c] = poly_eval(x[c], p[c], th);
                        } else {
                            double b[7];
                            dense_weights(th, b);
                            for (int c = 0; c < 4; ++c) q[c] = flat_dense_x(x[c], u[c], dt, b);
                        }
                        return min_distance(sc, q[0], q[1], q[2], q[3]);
                    };
                    double lo = th_lo, hi = th_hi;
                    double clo = cond_at(lo), chi = cond_at(hi);
                    double th_star;
                    if (chi == 0.0) {
                        th_star = hi;
                    } else {
                        int side = 0;
                        for (int it = 0; it < 100; ++it) {
                            if (!(hi - lo > 4.440892098500626e-16 * hi)) break;
                            // Illinois-modified regula falsi, bisection when the proposal leaves the bracket
                            double mid = lo - clo * (hi - lo) / (chi - clo);
                            if (!(mid > lo && mid < hi) || (it % 3) == 2) mid = lo + 0.5 * (hi - lo);
                            if (!(mid > lo && mid < hi)) break;
                            const double cm = cond_at(mid);
                            if (cm == 0.0) { lo = mid; clo = cm; break; }
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
                        th_star = lo;
                    }
                    if (th_star == 1.0) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) fs[c] = y[c];
                    } else if (th_star > 0.0) {
                        if (METRIC != RTGR_MINKOWSKI) {
                            for (int c = 0; c < 4; ++c) fs[c] = poly_eval(x[c], p[c], th_star);
                            dense_u(u, A, dt, th_star, fs + 4);
                        } else {
                            double b[7];
                            dense_weights(th_star, b);
                            for (int c = 0; c < 4; ++c) fs[c] = flat_dense_x(x[c], u[c], dt, b);
                        }
                    }
                }
                double col[3];
                const int omin = classify_color(sc, fs, col);
                if (job.rgb_f64) { for (int c = 0; c < 3; ++c) job.rgb_f64[3 * pix + c] = col[c]; }
                if (job.rgb8) {
                    uint8_t* o = job.rgb8 + 3 * (int64_t(pj) * sc.ni + pi);
                    o[0] = quantize8(col[0]); o[1] = quantize8(col[1]); o[2] = quantize8(col[2]);
                }
                if (job.final_state) { for (int c = 0; c < 8; ++c) job.final_state[8 * pix + c] = fs[c]; }
                if (job.obj_id) job.obj_id[pix] = omin;
                if (job.status) job.status[pix] = status;
                if (job.nsteps) job.nsteps[pix] = nacc;
                mode = L_IDLE;
            }
        }
    }
}

}  // namespace rtgr
