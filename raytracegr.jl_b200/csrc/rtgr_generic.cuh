// Generic-metric path: the reference accepts ANY callable `metric(x)` returning the
// 4x4 matrix g_ab (src:298-331) and obtains the Christoffel symbols by forward-mode automatic
// differentiation through it.  Here a user supplies that function as CUDA C++ source at run time
// (rtgr_metric_compile, NVRTC); it is compiled together with THIS file and the same integrator,
// event and colouring code as the built-in metrics (rtgr_trace.cuh).
//
// What the user writes (T is `double` for make_canvas and `Dual` for the geodesic right-hand side):
//
//     template <class T>
//     __device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par) { ... }
//
// x = (t, x, y, z), every entry of g must be assigned, `par` are up to 16 doubles set with
// rtgr_metric_set_params.  Available on T: + - * / (with T, double or int on either side), unary -,
// sqrt, pow2, pow3, pow4, powi(x, n), abs, sin, cos, exp, log, atan, atan2, acos, asin, cbrt -- the
// operator set of the reference's Dual type (src:51-196).
//
// Restated from the reference: Dual rules src:59-196 (the 2-argument atan follows the mathematically
// correct rule; the reference's line :168 is mis-parenthesised and is not on any path it runs),
// dmetric src:302-313, christoffel src:321-331, geodesic src:358-365, make_canvas src:464-476.
#pragma once

namespace rtgr_ad {

using ::sqrt; using ::fabs; using ::sin; using ::cos; using ::exp; using ::log; using ::atan;
using ::atan2; using ::acos; using ::asin; using ::cbrt;

// Number of partial derivatives a dual carries: 4 = d/dx^0..3 (src:11-14 with DT = SVector{4,T}); 3 = d/dx^1..3
// for a metric its author declares STATIONARY (`#pragma rtgr stationary` in the source: g does not depend on
// x[0], so d/dx^0 is identically zero and is neither carried nor multiplied -- a fifth of the arithmetic).
#ifndef RTGR_AD_NPART
#define RTGR_AD_NPART 4
#endif
constexpr int NP = RTGR_AD_NPART;
constexpr int POFF = 4 - NP;     // partial c is the derivative along coordinate c + POFF

// 1/x and sqrt(x), 1/(2 sqrt(x)) for the chain rules below: the branch-free MUFU-seeded Newton forms of the
// built-in path (rtgr_core.cuh, about 1 ulp) instead of the IEEE division / square-root subroutines.
RTGR_HD double rcp(double x) { return rtgr::fast_rcp(x); }

// value + gradient   (src:11-14)
struct Dual {
    double v;
    double e[NP];
    RTGR_HD Dual() : v(0.0) { for (int c = 0; c < NP; ++c) e[c] = 0.0; }
    RTGR_HD Dual(double a) : v(a) { for (int c = 0; c < NP; ++c) e[c] = 0.0; }        // src:16-21
    RTGR_HD Dual(int a) : v(double(a)) { for (int c = 0; c < NP; ++c) e[c] = 0.0; }
    RTGR_HD Dual(double a, int seed) : v(a) { for (int c = 0; c < NP; ++c) e[c] = (c + POFF == seed) ? 1.0 : 0.0; }
};

#define RTGR_AD_EACH for (int c = 0; c < NP; ++c)

RTGR_HD Dual operator+(const Dual& a) { return a; }
RTGR_HD Dual operator-(const Dual& a) { Dual r; r.v = -a.v; RTGR_AD_EACH r.e[c] = -a.e[c]; return r; }      // src:54-57
RTGR_HD Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; RTGR_AD_EACH r.e[c] = a.e[c] + b.e[c]; return r; }   // src:59-61
RTGR_HD Dual operator+(const Dual& a, double b) { Dual r = a; r.v = a.v + b; return r; }
RTGR_HD Dual operator+(double a, const Dual& b) { Dual r = b; r.v = a + b.v; return r; }
RTGR_HD Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; RTGR_AD_EACH r.e[c] = a.e[c] - b.e[c]; return r; }   // src:75-77
RTGR_HD Dual operator-(const Dual& a, double b) { Dual r = a; r.v = a.v - b; return r; }
RTGR_HD Dual operator-(double a, const Dual& b) { Dual r; r.v = a - b.v; RTGR_AD_EACH r.e[c] = -b.e[c]; return r; }
RTGR_HD Dual operator*(const Dual& a, const Dual& b) {                                                         // src:91-93
    Dual r; r.v = a.v * b.v; RTGR_AD_EACH r.e[c] = a.e[c] * b.v + a.v * b.e[c]; return r;
}
RTGR_HD Dual operator*(const Dual& a, double b) { Dual r; r.v = a.v * b; RTGR_AD_EACH r.e[c] = a.e[c] * b; return r; }
RTGR_HD Dual operator*(double a, const Dual& b) { Dual r; r.v = a * b.v; RTGR_AD_EACH r.e[c] = a * b.e[c]; return r; }
RTGR_HD Dual inv(const Dual& a) {                                                                              // src:107-110
    const double i = rcp(a.v), m = -i * i; Dual r; r.v = i; RTGR_AD_EACH r.e[c] = m * a.e[c]; return r;
}
RTGR_HD Dual operator/(const Dual& a, const Dual& b) {                                                         // src:112-114
    const double i = rcp(b.v), q = a.v * i; Dual r; r.v = q; RTGR_AD_EACH r.e[c] = (a.e[c] - q * b.e[c]) * i; return r;
}
RTGR_HD Dual operator/(const Dual& a, double b) { const double i = rcp(b); Dual r; r.v = a.v * i; RTGR_AD_EACH r.e[c] = a.e[c] * i; return r; }
RTGR_HD Dual operator/(double a, const Dual& b) { return a * inv(b); }
// int on either side (src:100-105, :118-120)
RTGR_HD Dual operator+(const Dual& a, int b) { return a + double(b); }
RTGR_HD Dual operator+(int a, const Dual& b) { return double(a) + b; }
RTGR_HD Dual operator-(const Dual& a, int b) { return a - double(b); }
RTGR_HD Dual operator-(int a, const Dual& b) { return double(a) - b; }
RTGR_HD Dual operator*(const Dual& a, int b) { return a * double(b); }
RTGR_HD Dual operator*(int a, const Dual& b) { return double(a) * b; }
RTGR_HD Dual operator/(const Dual& a, int b) { return a / double(b); }
RTGR_HD Dual operator/(int a, const Dual& b) { return double(a) / b; }
RTGR_HD Dual& operator+=(Dual& a, const Dual& b) { a = a + b; return a; }
RTGR_HD Dual& operator-=(Dual& a, const Dual& b) { a = a - b; return a; }
RTGR_HD Dual& operator*=(Dual& a, const Dual& b) { a = a * b; return a; }
RTGR_HD Dual& operator/=(Dual& a, const Dual& b) { a = a / b; return a; }

// chain rule: f(a) with derivative fp = f'(a.v)
RTGR_HD Dual chain(const Dual& a, double f, double fp) { Dual r; r.v = f; RTGR_AD_EACH r.e[c] = fp * a.e[c]; return r; }

RTGR_HD Dual pow2(const Dual& a) { return a * a; }                  // literal_pow src:134
RTGR_HD Dual pow3(const Dual& a) { return a * a * a; }              // src:135
RTGR_HD Dual pow4(const Dual& a) { return pow2(pow2(a)); }          // src:136
RTGR_HD double pow2(double a) { return a * a; }
RTGR_HD double pow3(double a) { return a * a * a; }
RTGR_HD double pow4(double a) { return pow2(pow2(a)); }
RTGR_HD double powi(double a, int n) {
    double r = 1.0, b = (n < 0) ? 1.0 / a : a;
    for (int k = (n < 0 ? -n : n); k > 0; k >>= 1) { if (k & 1) r *= b; b *= b; }
    return r;
}
RTGR_HD Dual powi(const Dual& a, int n) {                           // src:138-141
    if (n == 0) return Dual(1.0);
    return chain(a, powi(a.v, n), double(n) * powi(a.v, n - 1));
}
RTGR_HD Dual sqrt(const Dual& a) { double r; const double h = rtgr::fast_rsqrt_half(a.v, &r); return chain(a, r, h); }   // src:193-196
RTGR_HD Dual abs(const Dual& a) { return chain(a, fabs(a.v), a.v < 0.0 ? -1.0 : 1.0); }                      // src:150-152
RTGR_HD double abs(double a) { return fabs(a); }
RTGR_HD Dual sin(const Dual& a) { return chain(a, sin(a.v), cos(a.v)); }                                     // src:189-191
RTGR_HD Dual cos(const Dual& a) { return chain(a, cos(a.v), -sin(a.v)); }                                    // src:176-178
RTGR_HD Dual exp(const Dual& a) { const double r = exp(a.v); return chain(a, r, r); }                        // src:180-183
RTGR_HD Dual log(const Dual& a) { return chain(a, log(a.v), rcp(a.v)); }                                    // src:185-187
RTGR_HD Dual atan(const Dual& a) { return chain(a, atan(a.v), rcp(1.0 + a.v * a.v)); }                    // src:162-164
RTGR_HD Dual acos(const Dual& a) { return chain(a, acos(a.v), -1.0 / sqrt(1.0 - a.v * a.v)); }               // src:154-156
RTGR_HD Dual asin(const Dual& a) { return chain(a, asin(a.v), 1.0 / sqrt(1.0 - a.v * a.v)); }                // src:158-160
RTGR_HD Dual cbrt(const Dual& a) { const double r = cbrt(a.v); return chain(a, r, r / (3.0 * a.v)); }        // src:171-174
RTGR_HD Dual atan2(const Dual& y, const Dual& x) {                  // d atan(y/x) = (x dy - y dx)/(x^2 + y^2)
    const double i = rcp(x.v * x.v + y.v * y.v);
    Dual r; r.v = atan2(y.v, x.v); RTGR_AD_EACH r.e[c] = (x.v * y.e[c] - y.v * x.e[c]) * i; return r;
}
#undef RTGR_AD_EACH

// ---- the user's function (defined by the source handed to rtgr_metric_compile) ----------------
template <class T>
__device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par);
#ifdef RTGR_USER_KS_FORM
// ... or, for a metric of Kerr-Schild form  g_ab = eta_ab + f k_a k_b  (eta = diag(-1,1,1,1); any f and any k, null or
// not), the scalar and the covector alone:
template <class T>
__device__ void rtgr_user_kerr_schild(const T x[4], T& f, T k[4], const double* par);
#endif

// Closed-form inverse of a SYMMETRIC 4x4 matrix through 2x2 minors (what StaticArrays' inv does for 4x4: src:323, :470).
RTGR_HD void inverse4(const double m[4][4], double inv[4][4]) {
    const double s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1], s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
    const double s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3], s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    const double s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3], s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
    const double c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const double c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const double c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const double id = rcp(s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    // m is symmetric (a metric), hence so is its inverse: ten entries are formed, six mirrored
    inv[0][0] = (m[1][1] * c5 - m[1][2] * c4 + m[1][3] * c3) * id;
    inv[0][1] = inv[1][0] = (-m[0][1] * c5 + m[0][2] * c4 - m[0][3] * c3) * id;
    inv[0][2] = inv[2][0] = (m[3][1] * s5 - m[3][2] * s4 + m[3][3] * s3) * id;
    inv[0][3] = inv[3][0] = (-m[2][1] * s5 + m[2][2] * s4 - m[2][3] * s3) * id;
    inv[1][1] = (m[0][0] * c5 - m[0][2] * c2 + m[0][3] * c1) * id;
    inv[1][2] = inv[2][1] = (-m[3][0] * s5 + m[3][2] * s2 - m[3][3] * s1) * id;
    inv[1][3] = inv[3][1] = (m[2][0] * s5 - m[2][2] * s2 + m[2][3] * s1) * id;
    inv[2][2] = (m[3][0] * s4 - m[3][1] * s2 + m[3][3] * s0) * id;
    inv[2][3] = inv[3][2] = (-m[2][0] * s4 + m[2][1] * s2 - m[2][3] * s0) * id;
    inv[3][3] = (m[2][0] * s3 - m[2][1] * s1 + m[2][2] * s0) * id;
}

// geodesic acceleration  A^a = -Gamma^a_bc u^b u^c  for the user's metric (src:302-331, :358-365):
// dmetric by four seeded duals, then with  w_d = Gamma_{d,bc} u^b u^c = dg[d][b][c] u^b u^c - (1/2) dg[b][c][d] u^b u^c
// (the two "symmetric" terms of src:324-325 coincide under the contraction),  A^a = -g^{ad} w_d.
// directional derivative  u^c d_c q  of a dual (u = the 4-velocity)
RTGR_HD double along(const Dual& q, const double* u) {
    double t = 0.0;
    for (int c = 0; c < NP; ++c) t = fma(q.e[c], u[c + POFF], t);
    return t;
}

#ifdef RTGR_USER_KS_FORM
// Kerr-Schild form: with K = k_b u^b (u held fixed under the derivatives) and D = u^c d_c,
//   Gamma_{d,bc} u^b u^c = D(f K k_d) - (1/2) d_d (f K^2)
//                        = (Df K + f DK) k_d + f K Dk_d - (1/2) K^2 d_d f - f K u^a d_d k_a,
// raised with the Sherman-Morrison inverse  g^ad = eta^ad - f l^a l^d / (1 + f k.l),  l = eta k  (exact for any k).
// The 4x4 matrix, its ten derivative sets and the cofactor inverse of the general path are never formed.
__device__ __noinline__ void user_accel(const double* par, const double y[8], double A[4]) {
    Dual xd[4], f, k[4];
    for (int c = 0; c < 4; ++c) xd[c] = Dual(y[c], c);
    rtgr_user_kerr_schild<Dual>(xd, f, k, par);
    const double* u = y + 4;
    const double K = k[0].v * u[0] + k[1].v * u[1] + k[2].v * u[2] + k[3].v * u[3];
    const double Df = along(f, u);
    double Dk[4], DK = 0.0;
    for (int a = 0; a < 4; ++a) { Dk[a] = along(k[a], u); DK = fma(u[a], Dk[a], DK); }
    const double P = fma(Df, K, f.v * DK), Q = f.v * K, hK2 = 0.5 * K * K;
    double w[4];
    for (int d = 0; d < 4; ++d) {
        w[d] = fma(P, k[d].v, Q * Dk[d]);
        if (d >= POFF) {
            const int c = d - POFF;
            const double gK = k[0].e[c] * u[0] + k[1].e[c] * u[1] + k[2].e[c] * u[2] + k[3].e[c] * u[3];   // d_d K
            w[d] = fma(-hK2, f.e[c], fma(-Q, gK, w[d]));
        }
    }
    const double l[4] = {-k[0].v, k[1].v, k[2].v, k[3].v};
    const double kl = fma(k[0].v, l[0], fma(k[1].v, l[1], fma(k[2].v, l[2], k[3].v * l[3])));
    const double lw = fma(l[0], w[0], fma(l[1], w[1], fma(l[2], w[2], l[3] * w[3])));
    const double S = f.v * lw * rcp(fma(f.v, kl, 1.0));
    A[0] = fma(l[0], S, w[0]);                  // -(eta^00 w_0 - l^0 S) = w_0 + l^0 S
    for (int a = 1; a < 4; ++a) A[a] = fma(l[a], S, -w[a]);
}
#else
// geodesic acceleration  A^a = -Gamma^a_bc u^b u^c  for the user's metric (src:302-331, :358-365):
// dmetric by seeded duals, then with  w_d = Gamma_{d,bc} u^b u^c = dg[d][b][c] u^b u^c - (1/2) dg[b][c][d] u^b u^c
// (the two "symmetric" terms of src:324-325 coincide under the contraction),  A^a = -g^{ad} w_d.
__device__ __noinline__ void user_accel(const double* par, const double y[8], double A[4]) {
    Dual xd[4], g[4][4];
    for (int c = 0; c < 4; ++c) xd[c] = Dual(y[c], c);                // src:305-308
    rtgr_user_metric<Dual>(xd, g, par);                               // src:309
    const double* u = y + 4;
    // A metric tensor is symmetric: only g[a][b] with a <= b is read below (the reference evaluates and
    // sums both halves, which are equal), so the compiler drops the other six entries of the user's
    // function and their derivatives altogether.
    double G[4][4], gu[4][4], w[4] = {0.0, 0.0, 0.0, 0.0}, h[4] = {0.0, 0.0, 0.0, 0.0};
    for (int a = 0; a < 4; ++a)
        for (int b = a; b < 4; ++b) {
            const Dual& gab = g[a][b];
            G[a][b] = G[b][a] = gab.v;                                // src:310
            const double t = along(gab, u);                           // dg[a][b][c] u^c
            w[a] = fma(t, u[b], w[a]);
            if (b != a) w[b] = fma(t, u[a], w[b]);
            const double uab = (b != a ? 2.0 : 1.0) * u[a] * u[b];    // dg[a][b][d] u^a u^b, both halves
            for (int d = 0; d < NP; ++d) h[d + POFF] = fma(gab.e[d], uab, h[d + POFF]);
        }
    inverse4(G, gu);                                                  // src:323
    for (int d = 0; d < 4; ++d) w[d] = fma(-0.5, h[d], w[d]);
    for (int a = 0; a < 4; ++a)
        A[a] = -(gu[a][0] * w[0] + gu[a][1] * w[1] + gu[a][2] * w[2] + gu[a][3] * w[3]);   // src:326-330, :361-363
}
#endif

// make_canvas for one pixel with the user's metric (src:464-476); i, j are 0-based here
__device__ __noinline__ void user_canvas_pixel(const double* par, const double cam_pos[4], const double cam_wx[4],
                                               const double cam_wy[4], const double cam_n[4], int ni, int nj,
                                               int i, int j, double x[4], double u[4]) {
    const double dx = (double(i + 1) - 0.5) / double(ni) - 0.5;       // src:465-466
    const double dy = (double(j + 1) - 0.5) / double(nj) - 0.5;
    double n[4];
    for (int c = 0; c < 4; ++c) {
        x[c] = cam_pos[c] + dx * cam_wx[c] + dy * cam_wy[c];          // src:467
        n[c] = cam_n[c] + dx * cam_wx[c] + dy * cam_wy[c];            // src:468
    }
    double G[4][4], gu[4][4];
    rtgr_user_metric<double>(x, G, par);                              // src:469
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < a; ++b) G[a][b] = G[b][a];                // symmetric: the upper triangle is used
    inverse4(G, gu);                                                  // src:470
    double t[4], t2 = 0.0, n2 = 0.0;
    for (int a = 0; a < 4; ++a) t[a] = gu[a][0];                      // src:471: gu * (1,0,0,0)
    for (int a = 0; a < 4; ++a) {
        double gt = 0.0, gn = 0.0;
        for (int b = 0; b < 4; ++b) { gt += G[a][b] * t[b]; gn += G[a][b] * n[b]; }
        t2 += t[a] * gt; n2 += n[a] * gn;
    }
    const double st = sqrt(-t2), sn = sqrt(n2), s2 = sqrt(2.0);       // src:472-474
    for (int a = 0; a < 4; ++a) u[a] = (t[a] / st + n[a] / sn) / s2;
}

}  // namespace rtgr_ad
