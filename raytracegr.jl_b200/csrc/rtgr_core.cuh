// Per-ray arithmetic of the geodesic tracer: Kerr-Schild acceleration with hand-coded
// forward-mode derivatives, Tsit5 stage algebra in second-order form, error norm, PI step
// controller, event detection on the dense output, root-find, classification + colouring.
//
// Everything here is a per-lane, side-effect-free function on registers, written so that it
// compiles both for sm_100a (the product) and, through tests/host_shim.cpp, for the host (unit
// tests of the arithmetic on machines without a GPU -- test scaffolding only, never shipped).
//
// Reference semantics restated (src = RayTraceGR.jl/src/RayTraceGR.jl):
//   kerr_schild src:274-294, dmetric src:302-313, christoffel src:321-331, geodesic src:358-370,
//   distance/objcolor src:399-428, min_distance src:433-441, make_canvas src:464-476,
//   trace_rays src:485-533 (with the OrdinaryDiffEq/DiffEqBase behaviour of SURVEY.md appendix A).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/raytracegr_cuda.h"
#include "tsit5_tables.h"

#ifdef __CUDACC__
#define RTGR_HD __host__ __device__ __forceinline__
#else
#define RTGR_HD inline
#endif

// Strictly ordered, never-contracted multiply/add (used only by the Minkowski path, whose error
// estimate is pure rounding noise and therefore has to follow one fixed operation order).
#ifdef __CUDA_ARCH__
#define RTGR_MUL(a, b) __dmul_rn((a), (b))
#define RTGR_ADD(a, b) __dadd_rn((a), (b))
#else
#define RTGR_MUL(a, b) ((a) * (b)) /* host shim is compiled with -ffp-contract=off */
#define RTGR_ADD(a, b) ((a) + (b))
#endif

#ifdef __CUDACC__
#define RTGR_NOINLINE __host__ __device__ __noinline__
#else
#define RTGR_NOINLINE __attribute__((noinline))
#endif

namespace rtgr {

constexpr int MAX_INTERP = 32;

struct Vec4 { double v[4]; };
struct Vec8 { double v[8]; };

// ---------------------------------------------------------------------------------------------
// Branch-free FP64 reciprocal / reciprocal square root for the hot path.  The operands there are
// ordinary well-scaled numbers (radii, tolerances-scaled magnitudes), so the special-case slow
// paths of the IEEE division/sqrt sequences (a compare, a branch and a subroutine call each) are
// dead weight in the instruction stream.  MUFU seed (about 20 bits) + two Newton steps: ~1 ulp.
// Zero, negative (for rsqrt) or non-finite input ends up NaN/Inf, which the integrator catches as
// a NaN error estimate (status NONFINITE).
// ---------------------------------------------------------------------------------------------
RTGR_HD double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
    // seed y0 = (1/x)(1 - e), |e| <~ 2^-20;  1/x = y0 (1 + e + e^2 + O(e^3)):  three DFMAs, e^3 < 1/16 ulp
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
#else
    return 1.0 / x;
#endif
}
// One Newton step only (relative error ~1e-12): used where the quotient merely scales the error
// norm that drives the step-size controller.
RTGR_HD double fast_rcp_1nr(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}
// returns h = 1/(2 sqrt(x)); *root receives sqrt(x).  Coupled (Goldschmidt) iteration on
// g -> sqrt(x), h -> 1/(2 sqrt(x)):  r = 1/2 - g h;  g += g r;  h += h r  (the error squares each round): two
// rounds from the ~20-bit MUFU seed, 8 FP64 instructions for both results (about 1 ulp each); the factor 1/2 is
// what the callers want anyway (d sqrt = dx/(2 sqrt)).
RTGR_HD double fast_rsqrt_half(double x, double* root) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);                       // g = sqrt(x)(1+d), h = (1+d)/(2 sqrt(x)), d ~ 2^-39, the SAME d in both
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);                    // = -d - d^2/2
    *root = fma(g, r, g);                   // sqrt(x)(1 - 3d^2/2)
    return fma(h, r, h);                    // (1 - 3d^2/2)/(2 sqrt(x))
#else
    const double r = sqrt(x);
    *root = r;
    return 0.5 / r;
#endif
}
// returns 1/sqrt(x); *root receives sqrt(x)   (rare paths)
RTGR_HD double fast_rsqrt(double x, double* root) { return 2.0 * fast_rsqrt_half(x, root); }

// Scene + solver constants, flattened for __constant__ memory.
struct SceneConst {
    double M, a, a2, twoM, twoa2;     // twoa2 = 2 a^2
    double lambda0, lambda1, reltol, abstol, hit_threshold, dtmax;
    int32_t interp_points, maxiters, n_objs, metric;
    int32_t t1_half_hi, t1_quarter_hi;   // high words of lambda1/2 and lambda1/4 (see step_far_from_end)
    double theta[MAX_INTERP];  // theta[i] = i/(interp_points-1)
    int32_t kind[RTGR_MAX_OBJECTS];
    double sgn[RTGR_MAX_OBJECTS];   // sign(radius)
    double cx[RTGR_MAX_OBJECTS], cy[RTGR_MAX_OBJECTS], cz[RTGR_MAX_OBJECTS];
    double R2[RTGR_MAX_OBJECTS];    // radius^2
    double Rabs[RTGR_MAX_OBJECTS];  // |radius|
    double time[RTGR_MAX_OBJECTS];  // plane time
    // every distance function in one branch-free form (hot path):
    //   d_o(p) = qa*(px^2+py^2+pz^2) + qb0*pt + qb1*px + qb2*py + qb3*pz + qc
    double qa[RTGR_MAX_OBJECTS], qb0[RTGR_MAX_OBJECTS], qb1[RTGR_MAX_OBJECTS], qb2[RTGR_MAX_OBJECTS],
           qb3[RTGR_MAX_OBJECTS], qc[RTGR_MAX_OBJECTS];
    // event-filter margins: the curve may be `dev` (per component) off the chord, which can lower
    // d_o by at most mA*dev + mB*dev^2 (plane: dev; sphere: 2|R|*sqrt(3)dev + 3dev^2)
    double mA[RTGR_MAX_OBJECTS], mB[RTGR_MAX_OBJECTS];
    double qa_pos_max, mA_max, mB_max;   // maxima over the objects (coarse filter)
    double qa_pos_max_q;                 // qa_pos_max / 4
    double inv_nobj;                // 1/length(objs) is NOT used (division kept); n as double:
    double nobj_d;
    // camera (render mode)
    double cam_pos[4], cam_wx[4], cam_wy[4], cam_n[4];
    int32_t ni, nj;
    // parameters of a user-supplied metric (rtgr_metric_set_params)
    double user_par[16];
};

// template value of METRIC for a run-time compiled user metric (RTGR_MINKOWSKI = 0, RTGR_KERR_SCHILD = 1)
constexpr int METRIC_USER = 2;

}  // namespace rtgr
#ifdef RTGR_USER_METRIC
#include "rtgr_generic.cuh"
#endif
namespace rtgr {

// ---------------------------------------------------------------------------------------------
// Objects: distance (src:399-401, :415-419) and min_distance (src:433-441) at position p[0..3].
// ---------------------------------------------------------------------------------------------
RTGR_HD double obj_distance(const SceneConst& sc, int o, double pt, double px, double py, double pz) {
    if (sc.kind[o] == RTGR_PLANE) return pt - sc.time[o];
    const double dx = px - sc.cx[o], dy = py - sc.cy[o], dz = pz - sc.cz[o];
    return sc.sgn[o] * (dx * dx + dy * dy + dz * dz - sc.R2[o]);
}

RTGR_HD double min_distance(const SceneConst& sc, double pt, double px, double py, double pz) {
    double dmin = INFINITY;
#pragma unroll 1
    for (int o = 0; o < sc.n_objs; ++o) dmin = fmin(dmin, obj_distance(sc, o, pt, px, py, pz));
    return dmin;
}

// Branch-free forms used on the per-step path (same functions, expanded: sign(R)(|p-c|^2 - R^2) =
// sign(R)|p|^2 - 2 sign(R) c.p + sign(R)(|c|^2 - R^2); a plane has qa = 0).
RTGR_HD double obj_distance_q(const SceneConst& sc, int o, double n2, double pt, double px, double py, double pz) {
    return fma(sc.qa[o], n2, fma(sc.qb0[o], pt, fma(sc.qb1[o], px, fma(sc.qb2[o], py, fma(sc.qb3[o], pz, sc.qc[o])))));
}
RTGR_HD double min_distance_q(const SceneConst& sc, double pt, double px, double py, double pz) {
    const double n2 = fma(px, px, fma(py, py, pz * pz));
    double dmin = INFINITY;
#pragma unroll 1
    for (int o = 0; o < sc.n_objs; ++o) dmin = fmin(dmin, obj_distance_q(sc, o, n2, pt, px, py, pz));
    return dmin;
}

// The per-step form: up to four objects are evaluated as four INDEPENDENT chains in straight-line
// code with constant-bank coefficients (unused slots are padded with qc = +inf by
// build_scene_const), so that their latency overlaps and the compiler can interleave them with
// the error estimate; a sequential loop over the objects was the slowest stretch of the step
// (15 dependent DFMAs behind indexed constant loads).  Same value as min_distance_q.
RTGR_HD double min_distance_q4(const SceneConst& sc, double pt, double px, double py, double pz) {
    if (sc.n_objs > 4) return min_distance_q(sc, pt, px, py, pz);
    const double n2 = fma(px, px, fma(py, py, pz * pz));
    double d[4];
#pragma unroll
    for (int o = 0; o < 3; ++o) d[o] = obj_distance_q(sc, o, n2, pt, px, py, pz);
    const double m3 = fmin(fmin(d[0], d[1]), d[2]);
    if (sc.n_objs <= 3) return m3;      // the reference's scenes: caelum, frustum, sphere (src:546-549, :582-585)
    d[3] = obj_distance_q(sc, 3, n2, pt, px, py, pz);
    return fmin(m3, d[3]);
}

// What the per-step fast path needs of that minimum: "are ALL distances above a non-negative bound?".  The signed
// high words answer it (a negative distance has a negative high word; positive doubles order like their high
// words), so the minimum is taken over 32-bit integers: one VIMNMX3 instead of two FP64 compare-and-select
// sequences.  The exact minimum is recomputed by min_distance_q4 on the rare steps that need its value.
RTGR_HD int32_t hi_word_signed(double v) {
#ifdef __CUDA_ARCH__
    return __double2hiint(v);
#else
    int64_t b; memcpy(&b, &v, 8); return int32_t(b >> 32);
#endif
}
RTGR_HD int32_t imin3(int32_t a, int32_t b, int32_t c) { const int32_t m = a < b ? a : b; return m < c ? m : c; }
RTGR_HD int32_t min_distance_q4_hi(const SceneConst& sc, double pt, double px, double py, double pz) {
    if (sc.n_objs > 4) return hi_word_signed(min_distance_q(sc, pt, px, py, pz));
    const double n2 = fma(px, px, fma(py, py, pz * pz));
    int32_t h[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) h[o] = hi_word_signed(obj_distance_q(sc, o, n2, pt, px, py, pz));
    const int32_t m3 = imin3(h[0], h[1], h[2]);
    if (sc.n_objs <= 3) return m3;      // the reference's scenes: caelum, frustum, sphere (src:546-549, :582-585)
    const int32_t h3 = hi_word_signed(obj_distance_q(sc, 3, n2, pt, px, py, pz));
    return m3 < h3 ? m3 : h3;
}

// ---------------------------------------------------------------------------------------------
// Kerr-Schild geodesic acceleration  A^a = -Gamma^a_bc u^b u^c  (src:358-365 through :274-331),
// evaluated without ever forming g_ab,c or Gamma:
//
//   g_ab = eta_ab + f k_a k_b,  k = (1,k1,k2,k3),  f and k functions of (x,y,z) only.
//   Gamma_{d,bc} u^b u^c = w_d - v_d/2,   w_d = u^c d_c (f K k_d),   v_d = d_d (f K^2),
//   K = k_b u^b (u held fixed under the derivatives),
//   g^ad = eta^ad - f l^a l^d / (1 + f k.l),  l = eta k   (Sherman-Morrison: exact for ANY k,
//   so it also holds for the reference's radius formula, under which k is not null).
//
// Derivatives are hand-coded forward mode: one directional derivative D = u^c d_c and one
// spatial gradient d_i, both pushed through r(x,y,z) by the chain rule (d_t == 0: stationary).
// RFORM selects the radius line: as written at src:284 or the textbook one.
// ---------------------------------------------------------------------------------------------
template <int RFORM>
RTGR_HD void ks_accel(const SceneConst& sc, double x, double y, double z,
                      double ut, double ux, double uy, double uz, double A[4]) {
    const double a = sc.a, a2 = sc.a2;
    const double s = fma(z, z, fma(y, y, fma(x, x, -a2)));    // rho^2 - a^2
    const double h = 0.5 * s;
    const double az = a2 * z, az2 = az * z;
    double q;
    const double hq = fast_rsqrt_half(fma(h, h, az2), &q);   // 1/(2q)
    double r, Rs2, gzf;  // r; 2*dr/ds at fixed z; (dr/dz)/z = Rs2 + (dr/dz at fixed s)/z   (s = rho^2 - a^2)
    if (RFORM == RTGR_R_AS_WRITTEN) {
        double ss;      // NaN for rho < a: the ray is stopped (Julia would throw)
        const double hs = fast_rsqrt_half(s, &ss);        // 1/(2 sqrt s)
        r = fma(0.5, ss, q);
        Rs2 = fma(s, hq, hs);             // 2*(1/(4 sqrt s) + s/(4q))
        gzf = fma(sc.twoa2, hq, Rs2);     // + a^2 / q
    } else {
        const double i2r = fast_rsqrt_half(h + q, &r);    // 1/(2r)
        Rs2 = fma(s, hq, 1.0) * i2r;      // 2*(1/2 + s/(4q))/(2r)
        gzf = fma(sc.twoa2 * hq, i2r, Rs2);
    }
    // grad r
    const double gx = Rs2 * x, gy = Rs2 * y, gz = gzf * z;

    const double r2 = r * r, r3 = r2 * r;
    const double den = fma(r2, r2, az2);
    // 1/den, 1/r and 1/(r^2 + a^2) from ONE reciprocal of their product (one seed instead of three)
    const double ra = r2 + a2;
    const double rra = r * ra;
    const double ip = fast_rcp(rra * den);
    const double iden = ip * rra;
    const double ipd = ip * den;
    const double ir = ipd * ra;
    const double ira = ipd * r;
    const double r3i = r3 * iden;
    const double f = sc.twoM * r3i;                        // src:285
    // HALF of df/dr at fixed z and of -df/dz at fixed r (the factor 2 comes back below: Df = 2 Dfh; K^2/2 * F = K^2 * Fh)
    const double Frh = f * fma(-2.0, r3i, 1.5 * ir);
    const double Fzh = (f * iden) * az;                    // f a^2 z / (r^4 + a^2 z^2)
    const double rr = r * ira, aa = a * ira;               // r/(r^2+a^2), a/(r^2+a^2)
    const double k1 = fma(rr, x, aa * y);                  // src:287-289
    const double k2 = fma(rr, y, -(aa * x));
    const double k3 = z * ir;
    // d_i k_j = al_j g_i + B_ji,  B = ira*[[r,a,0],[-a,r,0],[0,0,(r^2+a^2)/r]];  al_3 = -k3/r is never formed:
    // it enters Au through -k3 (uz/r) and the force through P k3 + c1 al_3 = k3 (P - c1/r)
    const double tr = r + r;
    const double al1 = fma(-tr, k1, x) * ira;
    const double al2 = fma(-tr, k2, y) * ira;

    const double K = ut + k1 * ux + k2 * uy + k3 * uz;
    const double Dr = gx * ux + gy * uy + gz * uz;         // D r
    const double b3 = ir * uz;
    const double Au = fma(al1, ux, fma(al2, uy, -(k3 * b3)));
    // D k_j = al_j Dr + (B u)_j and d_i K = Au g_i + (B^T u)_i are never formed: only
    //   DK = u^j D k_j = Dr Au + u.B.u = Dr Au + rr (ux^2 + uy^2) + uz^2 / r      (B's antisymmetric part drops out)
    //   E_j = D k_j - d_j K = al_j Dr - Au g_j + 2 aa (uy, -ux, 0)_j              (only B's antisymmetric part stays)
    // enter, and E_j only through  Q E_j - (K^2/2) Fr g_j = c1 al_j - c2 g_j + c3 (uy, -ux, 0)_j.
    const double DK = fma(Dr, Au, fma(rr, fma(ux, ux, uy * uy), b3 * uz));
    const double Dfh = fma(Frh, Dr, -(Fzh * uz));          // Df / 2
    const double P = fma(Dfh, K + K, f * DK);              // Df K + f DK
    const double Q = f * K;
    const double KK = K * K;
    const double c1 = Q * Dr;
    const double c2 = fma(Q, Au, KK * Frh);
    const double c3h = Q * aa, c3 = c3h + c3h;
    // lower-index "force" F_d = w_d - v_d/2
    const double F0 = P;
    const double F1 = fma(c3, uy, fma(-c2, gx, fma(c1, al1, P * k1)));
    const double F2 = fma(-c3, ux, fma(-c2, gy, fma(c1, al2, P * k2)));
    const double F3 = fma(k3, fma(-c1, ir, P), fma(-c2, gz, KK * Fzh));
    // raise with g^ad and negate
    const double kk = k1 * k1 + k2 * k2 + k3 * k3;
    const double lF = k1 * F1 + k2 * F2 + k3 * F3 - F0;
    const double S = f * lF * fast_rcp(1.0 + f * (kk - 1.0));
    A[0] = F0 - S;
    A[1] = k1 * S - F1;
    A[2] = k2 * S - F2;
    A[3] = k3 * S - F3;
}

// The same acceleration for a == 0 (the reference's own scene: `a = 0  # T(0.8)`, src:276), where everything is a
// function of rho^2 alone:  k_j = x_j / r,  d_r k_j = -k_j / r,  f = 2M / r,  grad r = R' (x, y, z)  with
// r = sqrt(s)/2 + s/2 as written (s = rho^2; then k is not a unit vector) or r = rho (textbook), no z-dependence,
// no antisymmetric part.  The lower-index force collapses to F_j = x_j B.  59 FP64 instructions instead of 122
// and ONE square root + two reciprocals.  Internal template values RFORM + 2 select it (see variant_of).
constexpr int RFORM_A0 = 2;
template <int RF>
RTGR_HD void ks_accel_a0(const SceneConst& sc, double x, double y, double z,
                         double ut, double ux, double uy, double uz, double A[4]) {
    const double s = x * x + y * y + z * z;     // rho^2
    double ss;                                  // sqrt(s)
    const double hs = fast_rsqrt_half(s, &ss);  // 1/(2 sqrt s)
    double r, Rs2;                              // r; 2 dr/ds
    if (RF == RTGR_R_AS_WRITTEN) { r = fma(0.5, ss, 0.5 * s); Rs2 = 1.0 + hs; }
    else { r = ss; Rs2 = hs + hs; }
    const double ir = fast_rcp(r);
    const double f = sc.twoM * ir;              // src:285 with a = 0
    const double Fr = -(f * ir);                // df/dr
    const double xu = x * ux + y * uy + z * uz;
    const double ku = ir * xu;                  // k_j u^j
    const double K = ut + ku;
    const double Dr = Rs2 * xu;                 // D r
    const double Au = -(ir * ku);               // u^j d_r k_j
    const double usq = ux * ux + uy * uy + uz * uz;
    const double DK = fma(Dr, Au, ir * usq);
    const double P = fma(f, DK, (Fr * Dr) * K);
    const double Q = f * K;
    const double hK2 = 0.5 * K * K;
    const double c1 = Q * Dr;
    const double c2 = fma(Q, Au, hK2 * Fr);
    // F_j = P k_j + c1 d_r k_j - c2 grad_j r = x_j B
    const double B = fma(ir, fma(-c1, ir, P), -(c2 * Rs2));
    const double kk = (s * ir) * ir;
    const double lF = fma(ir * B, s, -P);       // k.F - F_0
    const double S = f * lF * fast_rcp(1.0 + f * (kk - 1.0));
    const double W = fma(ir, S, -B);            // A_j = k_j S - F_j = x_j (S/r - B)
    A[0] = P - S;
    A[1] = x * W;
    A[2] = y * W;
    A[3] = z * W;
}

// Kerr-Schild metric pieces at a point: f, k1..k3 (for make_canvas).
template <int RFORM>
RTGR_HD void ks_fk(const SceneConst& sc, double x, double y, double z, double& f, double k[3]) {
    const double a = sc.a, a2 = sc.a2;
    const double rho2 = x * x + y * y + z * z;
    const double s = rho2 - a2, h = 0.5 * s, az2 = a2 * z * z;
    double q, r;
    fast_rsqrt(az2 + h * h, &q);
    if ((RFORM & 1) == RTGR_R_AS_WRITTEN) { double ss; fast_rsqrt(s, &ss); r = 0.5 * ss + q; }   // (RFORM + 2: a == 0 variants)
    else fast_rsqrt(h + q, &r);
    const double r2 = r * r;
    f = sc.twoM * r2 * r * fast_rcp(r2 * r2 + az2);
    const double ira = fast_rcp(r2 + a2);
    k[0] = (r * x + a * y) * ira;
    k[1] = (r * y - a * x) * ira;
    k[2] = z * fast_rcp(r);
}

// ---------------------------------------------------------------------------------------------
// make_canvas for one pixel (src:464-476); i, j are 0-based here.  The Minkowski branch follows
// the reference's operation order exactly (its downstream error estimate is rounding noise, so
// the initial velocity has to be bit-reproducible); the Kerr-Schild branch uses the closed-form
// inverse metric.
// ---------------------------------------------------------------------------------------------
template <int METRIC, int RFORM>
RTGR_HD void canvas_pixel(const SceneConst& sc, int i, int j, double x[4], double u[4]) {
#ifdef RTGR_USER_METRIC
    if (METRIC == METRIC_USER) {
        rtgr_ad::user_canvas_pixel(sc.user_par, sc.cam_pos, sc.cam_wx, sc.cam_wy, sc.cam_n, sc.ni, sc.nj, i, j, x, u);
        return;
    }
#endif
    // (i - 1/2)/ni - 1/2 with the reference's 1-based i (src:465-466); IEEE division kept: the
    // pixel positions are inputs of everything else and must not depend on the code path
    const double dx = (double(i + 1) - 0.5) / double(sc.ni) - 0.5;
    const double dy = (double(j + 1) - 0.5) / double(sc.nj) - 0.5;
    double n[4];
    for (int c = 0; c < 4; ++c) {
        const double ox = RTGR_MUL(dx, sc.cam_wx[c]), oy = RTGR_MUL(dy, sc.cam_wy[c]);
        x[c] = RTGR_ADD(RTGR_ADD(sc.cam_pos[c], ox), oy);
        n[c] = RTGR_ADD(RTGR_ADD(sc.cam_n[c], ox), oy);
    }
    double t[4], st, sn;
    if (METRIC == RTGR_MINKOWSKI) {
        t[0] = -1.0; t[1] = t[2] = t[3] = 0.0;
        st = 1.0;
        double n2 = RTGR_MUL(n[0], -n[0]);
        n2 = RTGR_ADD(n2, RTGR_MUL(n[1], n[1]));
        n2 = RTGR_ADD(n2, RTGR_MUL(n[2], n[2]));
        n2 = RTGR_ADD(n2, RTGR_MUL(n[3], n[3]));
        sn = sqrt(n2);
    } else {
        double f, k[3];
        ks_fk<RFORM>(sc, x[1], x[2], x[3], f, k);
        const double kk = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
        const double w = f * fast_rcp(1.0 + f * (kk - 1.0));
        // t^a = g^{a0} = eta^{a0} - f l^a l^0/(1+f k.l),  l = (-1,k1,k2,k3)
        t[0] = -1.0 - w; t[1] = w * k[0]; t[2] = w * k[1]; t[3] = w * k[2];
        const double kt = t[0] + k[0] * t[1] + k[1] * t[2] + k[2] * t[3];
        const double kn = n[0] + k[0] * n[1] + k[1] * n[2] + k[2] * n[3];
        const double t2 = -t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3] + f * kt * kt;
        const double n2 = -n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3] + f * kn * kn;
        double dummy;
        const double ist = fast_rsqrt(-t2, &dummy), isn = fast_rsqrt(n2, &dummy);
        for (int c = 0; c < 4; ++c) u[c] = (t[c] * ist + n[c] * isn) * 0.7071067811865476;
        return;
    }
    const double s2 = sqrt(2.0);
    for (int c = 0; c < 4; ++c) u[c] = RTGR_ADD(t[c] / st, n[c] / sn) / s2;
}

// ---------------------------------------------------------------------------------------------
// Tsit5 in second-order form.  A[i] = acceleration of stage i+1 (4 components); (x,u) = state at
// the start of the step.  See gen_tables.py for the algebra.
// ---------------------------------------------------------------------------------------------

// The stage tables and the polynomial constants of the controller, as one block that lives in
// __constant__ memory on the device: with compile-time indices every coefficient becomes a
// constant-bank operand of its DFMA (no instruction spent on materialising it).
struct StageTab {
    double a[6][6];     // a[s-2][j]
    double abar[6][6];  // abar[s-2][i], zero-padded
    double c[6];
    double bt[7], btbar[6], btsum;
    double rbar[6][3];
    double logc[8];     // log(m) = 2s + s*z*(logc[0] + z*(logc[1] + ...)),  s = (m-1)/(m+1), z = s^2
    double expc[12];    // exp(r) = 1 + r*(expc[0] + r*(expc[1] + ...)),  expc[k] = 1/(k+1)!
    double ln2, inv_ln2, log_gamma, log_qoldinit, w_lo, w_hi, beta1, beta2, chord_dev;
};
constexpr StageTab make_stage_tab() {
    StageTab t{};
    for (int s = 0; s < 6; ++s) {
        for (int j = 0; j < 6; ++j) { t.a[s][j] = tab::a(s, j); t.abar[s][j] = (j < 5) ? tab::abar(s, j) : 0.0; }
        t.c[s] = tab::c(s);
    }
    for (int i = 0; i < 7; ++i) t.bt[i] = tab::bt(i);
    for (int i = 0; i < 6; ++i) { t.btbar[i] = tab::btbar(i); for (int m = 0; m < 3; ++m) t.rbar[i][m] = tab::rbar(i, m); }
    t.btsum = tab::btsum();
    for (int k = 0; k < 8; ++k) t.logc[k] = 2.0 / double(2 * k + 3);
    double f = 1.0;
    for (int k = 0; k < 12; ++k) { f *= double(k + 1); t.expc[k] = 1.0 / f; }
    t.ln2 = 0.6931471805599453; t.inv_ln2 = 1.4426950408889634;
    t.log_gamma = -0.10536051565782628;     // log(9/10)
    t.log_qoldinit = -9.210340371976182;    // log(1e-4)
    t.w_lo = -1.6094379124341003;           // log(qmin) = log(1/5)
    t.w_hi = 2.302585092994046;             // log(qmax) = log(10)
    t.beta1 = 7.0 / 50.0; t.beta2 = 2.0 / 25.0;
    t.chord_dev = tab::chord_dev_factor();
    return t;
}

// Stage state y_S (S = 2..7, compile-time) from the stored stage accelerations; y_7 is the
// candidate new state.  `acc.load(i, v)` returns the 4 acceleration components of stage i+1.
// NEED_T = false leaves y[0] (the time coordinate x^0 of the stage) untouched: the built-in metrics are
// stationary, so their right-hand side never reads it -- only the candidate state y_7 needs it.
template <int S, bool NEED_T, class Acc>
RTGR_HD void stage_state(const StageTab& T, const double x[4], const double u[4], const Acc& acc,
                         double dt, double dt2, double y[8]) {
    constexpr int C0 = NEED_T ? 0 : 1;
    double su[4], sx[4];
#pragma unroll
    for (int j = 0; j < S - 1; ++j) {
        double Aj[4];
        acc.load(j, Aj);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            su[c] = (j == 0) ? T.a[S - 2][0] * Aj[c] : fma(T.a[S - 2][j], Aj[c], su[c]);
            if (j < S - 2 && c >= C0) sx[c] = (j == 0) ? T.abar[S - 2][0] * Aj[c] : fma(T.abar[S - 2][j], Aj[c], sx[c]);
        }
    }
    const double dtc = dt * T.c[S - 2];      // x_S = x + (c_S dt) u + dt^2 sum_j abar_Sj A_j
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        y[4 + c] = fma(dt, su[c], u[c]);
        if (c < C0) continue;
        if (S == 2) y[c] = fma(dtc, u[c], x[c]);
        else y[c] = fma(dt2, sx[c], fma(dtc, u[c], x[c]));
    }
}

// Upper bound of |v| that costs no FP64 instruction: the exponent/high-mantissa word of a double
// orders magnitudes, so max over high words (+1 ulp of that word) bounds max |v|.
RTGR_HD uint32_t abs_hi_word(double v) {
#ifdef __CUDA_ARCH__
    return uint32_t(__double2hiint(v)) & 0x7fffffffu;
#else
    uint64_t b; memcpy(&b, &v, 8); return uint32_t(b >> 32) & 0x7fffffffu;
#endif
}
// The same bound through the FP32 min/max unit: read as a float, the high word of a double of magnitude in
// [2^-1015, 2^1017) is an ordinary normal float, and floats order like their bit patterns -- so the maximum of
// the |high words| is one FMNMX3 with |.| source modifiers per TWO new values (no sign-clearing instruction).
RTGR_HD float hi_word_as_float(double v) {
#ifdef __CUDA_ARCH__
    return __int_as_float(__double2hiint(v));
#else
    uint64_t b; memcpy(&b, &v, 8); const uint32_t h = uint32_t(b >> 32); float f; memcpy(&f, &h, 4); return f;
#endif
}
RTGR_HD uint32_t float_bits(float f) {
#ifdef __CUDA_ARCH__
    return uint32_t(__float_as_int(f));
#else
    uint32_t h; memcpy(&h, &f, 4); return h;
#endif
}
RTGR_HD uint32_t hi_word(double v) {
#ifdef __CUDA_ARCH__
    return uint32_t(__double2hiint(v));
#else
    uint64_t b; memcpy(&b, &v, 8); return uint32_t(b >> 32);
#endif
}
RTGR_HD double from_hi_word(uint32_t hi) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(int(hi), 0);
#else
    uint64_t b = uint64_t(hi) << 32; double v; memcpy(&v, &b, 8); return v;
#endif
}

// max(|a|, |b|) without touching the FP64 pipe: the magnitude bits of IEEE doubles order like
// unsigned integers.  (A NaN operand propagates instead of being dropped as fmax would; the callers'
// results are NaN in that case either way.)
#ifdef __CUDA_ARCH__
// Magnitude bits of a double.  The sign is cleared on the HIGH WORD inside inline PTX: written as a plain
// 64-bit AND the compiler recognises fabs() and emits DADD -RZ,|x| -- an instruction on the FP64 pipe,
// the very pipe these helpers are there to spare (16 of them per step attempt in the error norm).
__device__ __forceinline__ unsigned long long abs_bits(double v) {
    unsigned hi;
    asm("and.b32 %0, %1, 0x7fffffff;" : "=r"(hi) : "r"(__double2hiint(v)));
    return ((unsigned long long)hi << 32) | (unsigned)__double2loint(v);
}
#endif
// |v| without an FP64-pipe instruction
RTGR_HD double abs_int(double v) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)abs_bits(v));
#else
    return fabs(v);
#endif
}
RTGR_HD double max_abs(double a, double b) {
#ifdef __CUDA_ARCH__
    const unsigned long long ua = abs_bits(a), ub = abs_bits(b);
    return __longlong_as_double((long long)(ua > ub ? ua : ub));
#else
    const double fa = fabs(a), fb = fabs(b);
    return (fa != fa || fb != fb) ? (fa + fb) : fmax(fa, fb);
#endif
}

// Sign and order tests on the bit patterns (integer ALU instead of DSETP on the FP64 pipe, which is
// the kernel's bottleneck).  IEEE doubles of one sign order like integers.
RTGR_HD long long dbits(double v) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, 8); return b;
#endif
}
RTGR_HD double from_bits(long long b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, 8); return v;
#endif
}
RTGR_HD bool is_pos(double v) { return dbits(v) > 0; }                                     // v > 0
RTGR_HD bool is_neg(double v) { const long long b = dbits(v); return b < 0 && (b + b) != 0; }   // v < 0
RTGR_HD bool is_zero(double v) { const long long b = dbits(v); return (b + b) == 0; }        // v == +-0
// a > b for b >= 0 (any a; false for a <= 0)
RTGR_HD bool gt_nonneg(double a, double b) { return dbits(a) > dbits(b); }
// min(a, b) when at most one of them is negative
RTGR_HD double min_mixed(double a, double b) { const long long x = dbits(a), y = dbits(b); return from_bits(x < y ? x : y); }
// max(a, b) for a, b <= 0
RTGR_HD double max_nonpos(double a, double b) {
    const unsigned long long x = (unsigned long long)dbits(a), y = (unsigned long long)dbits(b);
    return from_bits((long long)(x < y ? x : y));
}
// v in [0, 1]   (v >= +0 or NaN)
RTGR_HD bool le_one_nonneg(double v) { return (unsigned long long)dbits(v) <= 0x3ff0000000000000ull; }
RTGR_HD bool is_nan_bits(double v) { return ((unsigned long long)dbits(v) & 0x7fffffffffffffffull) > 0x7ff0000000000000ull; }

// Embedded error estimate, scaled (A.2); returns the mean square of the residuals (= EEst^2).
// Also returns amax_hi: high-word bound of max |A_i,c| over stages 1..6 (for the event filter).
template <class Acc>
RTGR_HD double error_msq(const SceneConst& sc, const StageTab& T, const double x[4], const double u[4],
                         const Acc& acc, double dt, double dt2, const double y[8], uint32_t& amax_hi) {
    const double dtb = dt * T.btsum;
    double ex[4], eu[4];
    float am = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        double Ai[4];
        acc.load(i, Ai);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (i < 6) {
                ex[c] = (i == 0) ? T.btbar[0] * Ai[c] : fma(T.btbar[i], Ai[c], ex[c]);
                am = fmaxf(am, fabsf(hi_word_as_float(Ai[c])));
            }
            eu[c] = (i == 0) ? T.bt[0] * Ai[c] : fma(T.bt[i], Ai[c], eu[c]);
        }
    }
    amax_hi = float_bits(am);
    double sum = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double exc = fma(dt2, ex[c], dtb * u[c]);
        const double euc = dt * eu[c];
        // max(|uprev_i|, |u_i|) (A.2): one FP64 compare with |.| operands, the |.| of the winner folded into the DFMA
        const double mx = (fabs(x[c]) > fabs(y[c])) ? x[c] : y[c];
        const double mu = (fabs(u[c]) > fabs(y[4 + c])) ? u[c] : y[4 + c];
        const double scx = fma(fabs(mx), sc.reltol, sc.abstol);
        const double scu = fma(fabs(mu), sc.reltol, sc.abstol);
#ifdef RTGR_CONTROLLER_FP64
        const double rx = exc / scx, ru = euc / scu;     // letter-of-spec build: IEEE divisions (A.2)
#else
        const double rx = exc * fast_rcp_1nr(scx), ru = euc * fast_rcp_1nr(scu);
#endif
        sum = fma(rx, rx, sum);
        sum = fma(ru, ru, sum);
    }
    return sum * 0.125;
}

// Quartic coefficients of the dense output of the position components:
//   x_c(th) = x_c + th*(p1 + th*(p2 + th*(p3 + th*p4)))
template <class Acc>
RTGR_HD void dense_x_poly(const double u[4], const Acc& acc, double dt, double p[4][4]) {
    const double dt2 = dt * dt;
#pragma unroll
    for (int c = 0; c < 4; ++c) { p[c][0] = dt * u[c]; p[c][1] = 0.0; p[c][2] = 0.0; p[c][3] = 0.0; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double Ai[4];
        acc.load(i, Ai);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m = 0; m < 3; ++m) p[c][m + 1] = fma(tab::rbar(i, m), Ai[c], p[c][m + 1]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int m = 1; m < 4; ++m) p[c][m] *= dt2;
}
RTGR_HD double poly_eval(const double x0, const double p[4], double th) {
    return fma(th, fma(th, fma(th, fma(th, p[3], p[2]), p[1]), p[0]), x0);
}

// dense-output weights b_i(theta) (A.6), reference operation order
RTGR_HD void dense_weights(double th, double b[7]) {
    const double th2 = RTGR_MUL(th, th);
    b[0] = RTGR_MUL(th, RTGR_ADD(tab::r(0, 0), RTGR_MUL(th, RTGR_ADD(tab::r(0, 1), RTGR_MUL(th,
               RTGR_ADD(tab::r(0, 2), RTGR_MUL(th, tab::r(0, 3))))))));
#pragma unroll
    for (int i = 1; i < 7; ++i)
        b[i] = RTGR_MUL(th2, RTGR_ADD(tab::r(i, 1), RTGR_MUL(th, RTGR_ADD(tab::r(i, 2), RTGR_MUL(th, tab::r(i, 3))))));
}

// Velocity part of the dense output: u_c(th) = u_c + dt * sum_i b_i(th) A_i,c
template <class Acc>
RTGR_HD void dense_u(const double u[4], const Acc& acc, double dt, double th, double out[4]) {
    double b[7];
    dense_weights(th, b);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        double Ai[4];
        acc.load(i, Ai);
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] = fma(b[i], Ai[c], s[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = fma(dt, s[c], u[c]);
}

// ---------------------------------------------------------------------------------------------
// End-of-step distances and event filter.  Returns c1 = min_distance at the new point y and sets
// `clear` when NO object can be reached between the end points, so that the reference's interior
// dense-output samples (A.5) cannot change sign and may be skipped with an identical outcome.
// Along the chord p(th) = x + th*(y-x) every distance function is a parabola
//   d_o(th) = d0 + L th + Q th^2,  Q = qa |dxyz|^2,  L = d1 - d0 - Q,
// whose minimum over [0,1] is at an end point unless Q > 0 and the vertex lies inside.  The curve
// stays within `dev` (per component) of the chord, which lowers d_o by at most mA*dev + mB*dev^2.
// Conservative: clear == false only means "scan to be sure".
// ---------------------------------------------------------------------------------------------
// Coarse version of the same test from the minima alone: a parabola dips below its chord by at most
// Q/4, so  min_o d_o(th) >= min(c0, c1) - max_o(qa)^+ |dxyz|^2/4 - max_o margin.  Almost every step is
// far from every object and passes this; the per-object test below runs only for the rest.
// coarse_need returns that bound: the step is clear of every object when c0 > need and c1 > need.
RTGR_HD double coarse_need(const SceneConst& sc, const double x[4], const double y[8], double dev) {
    const double ex = y[1] - x[1], ey = y[2] - x[2], ez = y[3] - x[3];
    const double dd = fma(ex, ex, fma(ey, ey, ez * ez));
    // need = qa_max/4 |dxyz|^2 + mA_max dev + mB_max dev^2  (>= 0), in three FP64 instructions
    return fma(sc.qa_pos_max_q, dd, dev * fma(sc.mB_max, dev, sc.mA_max));
}

RTGR_HD double end_distances(const SceneConst& sc, const double x[4], const double y[8], double dev, bool& clear) {
    const double n0 = fma(x[1], x[1], fma(x[2], x[2], x[3] * x[3]));
    const double n1 = fma(y[1], y[1], fma(y[2], y[2], y[3] * y[3]));
    const double ex = y[1] - x[1], ey = y[2] - x[2], ez = y[3] - x[3];
    const double dd = fma(ex, ex, fma(ey, ey, ez * ez));
    const double dev2 = dev * dev;
    double c1 = INFINITY;
    bool ok = true;
#pragma unroll 1
    for (int o = 0; o < sc.n_objs; ++o) {
        const double d0 = obj_distance_q(sc, o, n0, x[0], x[1], x[2], x[3]);
        const double d1 = obj_distance_q(sc, o, n1, y[0], y[1], y[2], y[3]);
        c1 = fmin(c1, d1);
        const double margin = fma(sc.mA[o], dev, sc.mB[o] * dev2);
        const double Q = sc.qa[o] * dd;
        const double nL = d0 + Q - d1;                   // -L
        const double m0 = d0 - margin;
        bool good = fmin(d0, d1) > margin;
        // vertex inside (0,1): minimum d0 - L^2/(4Q) must clear the margin too
        if (Q > 0.0 && nL > 0.0 && nL < 2.0 * Q) good = good && (4.0 * Q * m0 > nL * nL * 1.0000000001);
        ok = ok && good;
    }
    clear = ok;
    return c1;
}

// ---- Minkowski: every stage slope is (u, 0); follow the reference order with k_i = u ---------
RTGR_HD double flat_sum_a7(double v) {  // ((((a71 v + a72 v) + a73 v) + ...) + a76 v)
    double acc = RTGR_MUL(tab::a(5, 0), v);
#pragma unroll
    for (int j = 1; j < 6; ++j) acc = RTGR_ADD(acc, RTGR_MUL(tab::a(5, j), v));
    return acc;
}
RTGR_HD double flat_sum_bt(double v) {
    double acc = RTGR_MUL(tab::bt(0), v);
#pragma unroll
    for (int j = 1; j < 7; ++j) acc = RTGR_ADD(acc, RTGR_MUL(tab::bt(j), v));
    return acc;
}
RTGR_HD double flat_dense_x(double x0, double v, double dt, const double b[7]) {
    double acc = RTGR_MUL(v, b[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) acc = RTGR_ADD(acc, RTGR_MUL(v, b[i]));
    return RTGR_ADD(x0, RTGR_MUL(dt, acc));
}
RTGR_HD double flat_error_msq(const SceneConst& sc, const double x[4], const double u[4], double dt,
                              const double y[8]) {
    double sum = 0.0;
    for (int c = 0; c < 4; ++c) {   // position residuals first (state order x0..x3,u0..u3); the
        const double e = RTGR_MUL(dt, flat_sum_bt(u[c]));   // velocity residuals are exactly zero
        const double scl = RTGR_ADD(sc.abstol, RTGR_MUL(fmax(fabs(x[c]), fabs(y[c])), sc.reltol));
        const double r = e / scl;
        sum = RTGR_ADD(sum, RTGR_MUL(r, r));
    }
    return sum / 8.0;
}

// ---------------------------------------------------------------------------------------------
// PI step-size controller (A.3) in log form: q = EEst^beta1 / qold^beta2 / gamma, clamped to
// [1/qmax, 1/qmin].  lE = log(EEst), lqold = log(qold).
// ---------------------------------------------------------------------------------------------
constexpr double BETA1 = 7.0 / 50.0, BETA2 = 2.0 / 25.0, GAMMA = 9.0 / 10.0;
constexpr double QMIN = 1.0 / 5.0, QMAX = 10.0;
constexpr double LOG_QOLDINIT = -9.210340371976182;  // log(1e-4)

// Natural log of a positive, normal double (no special cases): exponent split, then the atanh
// series in s = (m-1)/(m+1) with m in [sqrt(1/2), sqrt(2)).  Max error ~2 ulp.
RTGR_HD double log_pos(const StageTab& T, double v) {
#ifdef __CUDA_ARCH__
    int hi = __double2hiint(v);
    const int lo = __double2loint(v);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }   // m >= sqrt(2): halve
    const double m = __hiloint2double(hi, lo);
#else
    int e;
    double m = frexp(v, &e) * 2.0; e -= 1;                  // m in [1,2)
    if (m >= 1.4142135623730951) { m *= 0.5; e += 1; }
#endif
    const double f = m - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s;
    double p = T.logc[7];
#pragma unroll
    for (int k = 6; k >= 0; --k) p = fma(p, z, T.logc[k]);
    const double lm = fma(s * z, p, s + s);
    return fma(double(e), T.ln2, lm);
}

// exp(w) for |w| <= ~3 (the controller clamps first): n = rint(w/ln2), Taylor in r = w - n ln2.
RTGR_HD double exp_small(const StageTab& T, double w) {
    const double nd = rint(w * T.inv_ln2);
    const double r = fma(-nd, T.ln2, w);
    double p = T.expc[11];
#pragma unroll
    for (int k = 10; k >= 0; --k) p = fma(p, r, T.expc[k]);
    const double er = fma(p, r, 1.0);
#ifdef __CUDA_ARCH__
    const int n = int(nd);
    return er * __hiloint2double((n + 1023) << 20, 0);
#else
    return ldexp(er, int(nd));
#endif
}

// PI controller (A.3) in log form.  Returns 1/q = clamp(gamma * EEst^-beta1 * qold^beta2, 1/5, 10)
// (the step is multiplied by it on acceptance) and lE = log(EEst).
RTGR_HD double controller_inv_q(const StageTab& T, double msq, double lqold, double& lE) {
    // EEst == 0 (or underflow, or NaN): q = 1/qmax.   0x0010... = smallest normal double
    if ((unsigned long long)dbits(msq) - 0x0010000000000000ull >= 0x7fe0000000000000ull) { lE = -INFINITY; return QMAX; }
    lE = 0.5 * log_pos(T, msq);
    double w = fma(-T.beta1, lE, fma(T.beta2, lqold, T.log_gamma));
    // clamp to [w_lo, w_hi] (w_lo < 0 < w_hi) on the bit patterns
    if (dbits(w) > dbits(T.w_hi)) w = T.w_hi;
    else if ((unsigned long long)dbits(w) > (unsigned long long)dbits(T.w_lo)) w = T.w_lo;
    return exp_small(T, w);
}
// The same controller with the factor evaluated on the otherwise idle FP32/SFU pipes (Kerr-Schild
// path).  Only the step-size FACTOR is computed here: the accept/reject decision (EEst <= 1) and all
// state arithmetic stay FP64.  lg2/ex2.approx are good to ~2^-21, so dt differs from the FP64
// controller's by ~1e-7 relative -- that changes a step's truncation error (itself <= tol) by ~5e-7 of
// itself, far below FP64 rounding of the state.  Logs are base 2 and kept as float lane state.
// At the clamps the exact constants 10 and 1/5 are returned.
constexpr float LOG2_QOLDINIT_F = -13.287712379549449f;   // log2(1e-4)
constexpr float LOG2_GAMMA_F = -0.15200309344504997f;     // log2(9/10)
constexpr float W2_LO_F = -2.321928094887362f;            // log2(1/5)
constexpr float W2_HI_F = 3.321928094887362f;             // log2(10)
RTGR_HD float lg2_approx(float v) {
#ifdef __CUDA_ARCH__
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
    return log2f(v);
#endif
}
RTGR_HD float ex2_approx(float v) {
#ifdef __CUDA_ARCH__
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
    return exp2f(v);
#endif
}
RTGR_HD double controller_inv_q_fast(double msq, float lqold2, float& lE2) {
    // EEst^2 as a float: zero / underflow gives lg2 = -inf and the clamp q = 1/qmax (A.3: EEst == 0), overflow or
    // NaN the other clamp (a NaN estimate is rejected by the caller's own FP64 test before the factor is used)
    const float m = float(msq);
    lE2 = 0.5f * lg2_approx(m);                                                 // log2(EEst)
    const float w = fmaf(-float(BETA1), lE2, fmaf(float(BETA2), lqold2, LOG2_GAMMA_F));
    double q = double(ex2_approx(fminf(fmaxf(w, W2_LO_F), W2_HI_F)));           // in [1/5, 10] up to 2^-21
    if (w >= W2_HI_F) q = QMAX;                                                 // exact constants at the clamps
    if (!(w > W2_LO_F)) q = QMIN;
    return q;
}

RTGR_NOINLINE double reject_factor(double lE) {  // dt <- dt * this (rare: out of line)
    const double q11 = exp(BETA1 * lE);
    return 1.0 / fmin(1.0 / QMIN, q11 / GAMMA);
}

// ---------------------------------------------------------------------------------------------
// Classification + colouring (src:513-533).  Returns omin (0 = hit nothing).
// ---------------------------------------------------------------------------------------------
RTGR_HD double jl_mod1(double v) {  // Julia mod(v, 1) = v - floor(v) for |v| far below 2^52
    return v - floor(v);
}

RTGR_HD int classify_color(const SceneConst& sc, const double p[4], double col[3]) {
    int omin = 0;
    double dmin = sc.hit_threshold;
#pragma unroll 1
    for (int o = 0; o < sc.n_objs; ++o) {
        const double d = obj_distance(sc, o, p[0], p[1], p[2], p[3]);
        if (d < dmin) { omin = o + 1; dmin = d; }
    }
    if (omin == 0) { col[0] = 1.0; col[1] = 0.0; col[2] = 0.0; return 0; }
    const int o = omin - 1;
    if (sc.kind[o] == RTGR_PLANE) {
        col[0] = 0.0; col[1] = 0.5; col[2] = 0.0;
    } else {
        const double x = p[1] - sc.cx[o], y = p[2] - sc.cy[o], z = p[3] - sc.cz[o];
        const double r = sqrt(x * x + y * y + z * z);
        const double th = acos(z / r);
        const double ph = atan2(y, x);
        const double pi = 3.14159265358979323846;
        col[0] = jl_mod1(12.0 * th / pi);
        col[1] = jl_mod1(12.0 * ph / pi);
        col[2] = 1.0;
    }
    const double w = double(omin) / sc.nobj_d;
    col[0] *= w; col[1] *= w; col[2] *= w;
    return omin;
}

RTGR_HD uint8_t quantize8(double v) {  // Float64 -> N0f8: round(255 x), clamped
    v = fmin(1.0, fmax(0.0, v));
    return (uint8_t)rint(255.0 * v);
}

}  // namespace rtgr
