// =====================================================================================
// ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17, no dependencies) of the algorithm of RayTraceGR.jl's
// per-pixel geodesic ray trace, in the reference's own operation order ("as written").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (libraytracegr_cuda) never links or calls it.
//
// What it follows (src = /root/reference/src/RayTraceGR.jl):
//   Dual arithmetic            src:11-14, :59-120, :132-136, :193-196
//   minkowski / kerr_schild    src:262-264 / src:274-294 (incl. the radius line :284 as is)
//   dmetric / christoffel      src:302-313 / src:321-331
//   geodesic RHS               src:358-370
//   Plane / Sphere / distance / objcolor / min_distance   src:399-404, :415-428, :433-441
//   make_canvas                src:464-476
//   trace_rays                 src:485-533
// The ODE machinery trace_rays calls is third-party Julia code that is NOT in the
// reference tree (only pinned in Manifest.toml): OrdinaryDiffEq 5.38.3 (Tsit5 constant
// cache, PI controller, dense output, initial dt), DiffEqBase 6.35.2 (ContinuousCallback,
// error norm), Roots 1.0.1 (bracketing root-find), StaticArrays 0.12.3 (4x4 inverse).  Its
// published algorithm is restated here.
//
// Parity pin: the reference has no test that exercises this path (test/runtests.jl:65-79
// is commented out).  What pins the oracle are the two golden images the reference ships,
// scenes/sphere.png and scenes/sphere2.png (committed as tests/golden/*.npy): this oracle
// reproduces sphere2.png on 40000/40000 pixels and sphere.png on all but silhouette-edge
// pixels (tests/test_oracle.py).  The Julia code itself cannot be run in this image
// (no Julia), so bit-level agreement with the Julia solver is NOT claimed.
// =====================================================================================
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/raytracegr_cuda.h"

namespace {

constexpr int D = 4;  // src:254

// ------------------------------------------------------------------------------------
// Dual number with a 4-vector of partials (src:11-14).  Only the methods the metric code
// reaches are provided; each cites the rule it restates.
// ------------------------------------------------------------------------------------
template <class T>
struct Dual {
    T val;
    T eps[D];
    Dual() : val(0) { for (auto& e : eps) e = T(0); }
    Dual(T v) : val(v) { for (auto& e : eps) e = T(0); }  // src:16-21
    Dual(T v, int seed) : val(v) { for (int c = 0; c < D; ++c) eps[c] = (c == seed) ? T(1) : T(0); }
};

template <class T> Dual<T> operator+(const Dual<T>& x, const Dual<T>& y) {  // src:59-61
    Dual<T> r; r.val = x.val + y.val;
    for (int c = 0; c < D; ++c) r.eps[c] = x.eps[c] + y.eps[c];
    return r;
}
template <class T> Dual<T> operator-(const Dual<T>& x, const Dual<T>& y) {  // src:75-77
    Dual<T> r; r.val = x.val - y.val;
    for (int c = 0; c < D; ++c) r.eps[c] = x.eps[c] - y.eps[c];
    return r;
}
template <class T> Dual<T> operator-(const Dual<T>& x, T a) {  // src:78-80, :84-86
    Dual<T> r = x; r.val = x.val - a; return r;
}
template <class T> Dual<T> operator*(const Dual<T>& x, const Dual<T>& y) {  // src:91-93
    Dual<T> r; r.val = x.val * y.val;
    for (int c = 0; c < D; ++c) r.eps[c] = x.eps[c] * y.val + x.val * y.eps[c];
    return r;
}
template <class T> Dual<T> operator*(T a, const Dual<T>& x) {  // src:97-99, :103-105
    Dual<T> r; r.val = a * x.val;
    for (int c = 0; c < D; ++c) r.eps[c] = a * x.eps[c];
    return r;
}
template <class T> Dual<T> operator/(const Dual<T>& x, const Dual<T>& y) {  // src:112-114
    Dual<T> r; r.val = x.val / y.val;
    const T y2 = y.val * y.val;
    for (int c = 0; c < D; ++c) r.eps[c] = (x.eps[c] * y.val - x.val * y.eps[c]) / y2;
    return r;
}
template <class T> Dual<T> operator/(const Dual<T>& x, T a) {  // src:115-120
    Dual<T> r; r.val = x.val / a;
    for (int c = 0; c < D; ++c) r.eps[c] = x.eps[c] / a;
    return r;
}
template <class T> Dual<T> sqrt(const Dual<T>& x) {  // src:193-196
    using std::sqrt;
    Dual<T> r; r.val = sqrt(x.val);
    const T w = T(1) / (T(2) * r.val);
    for (int c = 0; c < D; ++c) r.eps[c] = w * x.eps[c];
    return r;
}
template <class T> Dual<T> pow2(const Dual<T>& x) { return x * x; }           // src:134
template <class T> Dual<T> pow3(const Dual<T>& x) { return x * x * x; }       // src:135
template <class T> Dual<T> pow4(const Dual<T>& x) { return pow2(pow2(x)); }   // src:136

// the same spellings for plain scalars, so one metric template serves both
inline float pow2(float x) { return x * x; }
inline float pow3(float x) { return x * x * x; }
inline float pow4(float x) { return pow2(pow2(x)); }
inline double pow2(double x) { return x * x; }
inline double pow3(double x) { return x * x * x; }
inline double pow4(double x) { return pow2(pow2(x)); }
inline long double pow2(long double x) { return x * x; }
inline long double pow3(long double x) { return x * x * x; }
inline long double pow4(long double x) { return pow2(pow2(x)); }

template <class S> struct scalar_of { using type = S; };
template <class T> struct scalar_of<Dual<T>> { using type = T; };

template <class S> using Mat4 = std::array<std::array<S, D>, D>;

struct MetricSpec {
    int kind;       // rtgr_metric_kind
    int r_formula;  // rtgr_r_formula
    double M, a;
};

// ------------------------------------------------------------------------------------
// Metrics.  S is either a plain scalar or a Dual over it.
// ------------------------------------------------------------------------------------
template <class S>
Mat4<S> minkowski(const std::array<S, D>&) {  // src:262-264
    using T = typename scalar_of<S>::type;
    Mat4<S> g;
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b) g[a][b] = S(T(a == b ? (a == 0 ? -1 : 1) : 0));
    return g;
}

template <class S>
Mat4<S> kerr_schild(const std::array<S, D>& xx, const MetricSpec& ms, bool* domain_ok) {  // src:274-294
    using T = typename scalar_of<S>::type;
    using std::sqrt;
    const T M = T(ms.M);  // src:275 (reference value 1)
    const T a = T(ms.a);  // src:276 (reference value 0)
    const S x = xx[1], y = xx[2], z = xx[3];
    Mat4<S> eta = minkowski<S>(xx);                       // src:282
    const S rho = sqrt(pow2(x) + pow2(y) + pow2(z));      // src:283
    const T a2 = a * a;
    const S h = (pow2(rho) - a2) / T(2);                  // (rho^2 - a^2)/2
    const S q = sqrt(a2 * pow2(z) + pow2(h));
    S r;
    if (ms.r_formula == RTGR_R_AS_WRITTEN) {
        r = sqrt(pow2(rho) - a2) / T(2) + q;              // src:284, literally
    } else {
        r = sqrt(h + q);                                  // textbook Kerr-Schild radius
    }
    const S f = (T(2) * M) * pow3(r) / (pow4(r) + a2 * pow2(z));   // src:285
    std::array<S, D> k;                                    // src:286-289
    k[0] = S(T(1));
    k[1] = (r * x + a * y) / (pow2(r) + S(a2));
    k[2] = (r * y - a * x) / (pow2(r) + S(a2));
    k[3] = z / r;
    Mat4<S> g;
    for (int p = 0; p < D; ++p)
        for (int s = 0; s < D; ++s) g[p][s] = eta[p][s] + f * k[p] * k[s];   // src:291
    if (domain_ok) {
        T rv;
        if constexpr (std::is_same<S, T>::value) rv = r; else rv = r.val;
        *domain_ok = !(rv != rv);
    }
    return g;
}

// A metric the reference does not ship, for the parity tests of the user-metric path (the reference's
// trace_rays accepts any callable, src:483): Schwarzschild in isotropic coordinates,
//   ds^2 = -((1 - m/2rho)/(1 + m/2rho))^2 dt^2 + (1 + m/2rho)^4 (dx^2 + dy^2 + dz^2),  m = ms.M.
// Evaluated through the same Dual arithmetic / dmetric / christoffel as any other metric.
constexpr int ORACLE_METRIC_SCHWARZSCHILD_ISOTROPIC = 1000;
template <class S>
Mat4<S> schwarzschild_isotropic(const std::array<S, D>& xx, const MetricSpec& ms) {
    using T = typename scalar_of<S>::type;
    using std::sqrt;
    const S one = S(T(1)), zero = S(T(0));
    const S rho = sqrt(pow2(xx[1]) + pow2(xx[2]) + pow2(xx[3]));
    const S w = S(T(ms.M) / T(2)) / rho;
    const S psi = one + w;
    Mat4<S> g;
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b) g[a][b] = zero;
    g[0][0] = zero - pow2((one - w) / psi);
    g[1][1] = g[2][2] = g[3][3] = pow4(psi);
    return g;
}

template <class S>
Mat4<S> eval_metric(const std::array<S, D>& x, const MetricSpec& ms, bool* ok = nullptr) {
    if (ok) *ok = true;
    if (ms.kind == RTGR_MINKOWSKI) return minkowski<S>(x);
    if (ms.kind == ORACLE_METRIC_SCHWARZSCHILD_ISOTROPIC) return schwarzschild_isotropic<S>(x, ms);
    return kerr_schild<S>(x, ms, ok);
}

// ------------------------------------------------------------------------------------
// dmetric (src:302-313): seed four duals, evaluate, split into g[a][b], dg[a][b][c].
// ------------------------------------------------------------------------------------
template <class T>
void dmetric(const MetricSpec& ms, const std::array<T, D>& x, Mat4<T>& g, T dg[D][D][D], bool* ok) {
    std::array<Dual<T>, D> xdx;
    for (int c = 0; c < D; ++c) xdx[c] = Dual<T>(x[c], c);   // src:305-308
    Mat4<Dual<T>> gdg = eval_metric<Dual<T>>(xdx, ms, ok);     // src:309
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b) {
            g[a][b] = gdg[a][b].val;                            // src:310
            for (int c = 0; c < D; ++c) dg[a][b][c] = gdg[a][b].eps[c];   // src:311
        }
}

// Closed-form 4x4 inverse, adjugate / determinant (what StaticArrays' inv does for 4x4,
// src:323 / src:470).
template <class T>
T det3(T a, T b, T c, T d, T e, T f, T g, T h, T i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}
template <class T>
Mat4<T> inverse4(const Mat4<T>& A, T* det_out = nullptr) {
    Mat4<T> C;  // cofactors
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            int r[3], c[3];
            for (int k = 0, n = 0; k < D; ++k) if (k != i) r[n++] = k;
            for (int k = 0, n = 0; k < D; ++k) if (k != j) c[n++] = k;
            T m = det3(A[r[0]][c[0]], A[r[0]][c[1]], A[r[0]][c[2]],
                       A[r[1]][c[0]], A[r[1]][c[1]], A[r[1]][c[2]],
                       A[r[2]][c[0]], A[r[2]][c[1]], A[r[2]][c[2]]);
            C[i][j] = ((i + j) & 1) ? -m : m;
        }
    T det = A[0][0] * C[0][0] + A[0][1] * C[0][1] + A[0][2] * C[0][2] + A[0][3] * C[0][3];
    if (det_out) *det_out = det;
    const T idet = T(1) / det;
    Mat4<T> B;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) B[i][j] = C[j][i] * idet;
    return B;
}

// ------------------------------------------------------------------------------------
// christoffel (src:321-331)
// ------------------------------------------------------------------------------------
template <class T>
void christoffel(const MetricSpec& ms, const std::array<T, D>& x, T Gam[D][D][D], bool* ok) {
    Mat4<T> g; T dg[D][D][D];
    dmetric<T>(ms, x, g, dg, ok);                       // src:322
    Mat4<T> gu = inverse4<T>(g);                        // src:323
    T Gl[D][D][D];
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b)
            for (int c = 0; c < D; ++c)
                Gl[a][b][c] = (dg[a][b][c] + dg[a][c][b] - dg[b][c][a]) / T(2);   // src:324-325
    for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b)
            for (int c = 0; c < D; ++c)
                Gam[a][b][c] = gu[a][0] * Gl[0][b][c] + gu[a][1] * Gl[1][b][c] +
                               gu[a][2] * Gl[2][b][c] + gu[a][3] * Gl[3][b][c];    // src:326-330
}

template <class T> using State = std::array<T, 2 * D>;   // r2s / s2r packing, src:345-352

// geodesic (src:358-370): xdot = u, udot^a = - sum_{x,y} Gamma[a,x,y] u^x u^y
template <class T>
State<T> geodesic(const State<T>& s, const MetricSpec& ms, bool* ok = nullptr) {
    State<T> out;
    std::array<T, D> x{s[0], s[1], s[2], s[3]};
    const T* u = &s[D];
    T Gam[D][D][D];
    christoffel<T>(ms, x, Gam, ok);
    for (int a = 0; a < D; ++a) {
        out[a] = u[a];                                   // src:360
        // Julia's sum over an SMatrix runs in column-major order: x fastest (src:361-363)
        T acc = T(0);
        bool first = true;
        for (int y = 0; y < D; ++y)
            for (int xx = 0; xx < D; ++xx) {
                T term = Gam[a][xx][y] * u[xx] * u[y];
                acc = first ? term : acc + term;
                first = false;
            }
        out[D + a] = -acc;
    }
    return out;
}

// ------------------------------------------------------------------------------------
// Objects (src:393-441)
// ------------------------------------------------------------------------------------
template <class T>
T obj_distance(const rtgr_object& o, const T* pos) {
    if (o.kind == RTGR_PLANE) return pos[0] - T(o.time);               // src:399-401
    const T R = T(o.radius);
    const T d1 = pos[1] - T(o.pos[1]), d2 = pos[2] - T(o.pos[2]), d3 = pos[3] - T(o.pos[3]);
    const T s = d1 * d1 + d2 * d2 + d3 * d3;
    const T sg = R > 0 ? T(1) : (R < 0 ? T(-1) : T(0));
    return sg * (s - R * R);                                           // src:415-419
}

template <class T>
T jl_mod1(T x) {  // Julia mod(x, 1) for floats
    using std::fmod;
    T r = fmod(x, T(1));
    if (r == 0) return std::fabs(r);
    if (r < 0) return r + T(1);
    return r;
}

template <class T>
void obj_color(const rtgr_object& o, const T* pos, T col[3]) {
    if (o.kind == RTGR_PLANE) { col[0] = 0; col[1] = T(1) / 2; col[2] = 0; return; }   // src:402-404
    using std::sqrt; using std::acos; using std::atan2;
    const T x = pos[1] - T(o.pos[1]), y = pos[2] - T(o.pos[2]), z = pos[3] - T(o.pos[3]);
    const T r = sqrt(x * x + y * y + z * z);
    const T th = acos(z / r);
    const T ph = atan2(y, x);
    const T pi = T(3.14159265358979323846264338327950288L);
    col[0] = jl_mod1(T(12) * th / pi);                                  // src:427
    col[1] = jl_mod1(T(12) * ph / pi);
    col[2] = T(1);
}

template <class T>
T min_distance(const rtgr_object* objs, int n, const T* pos) {   // src:433-441
    T dmin = std::numeric_limits<T>::infinity();
    for (int o = 0; o < n; ++o) dmin = std::min(dmin, obj_distance<T>(objs[o], pos));
    return dmin;
}

// ------------------------------------------------------------------------------------
// make_canvas, one pixel (src:464-476); i,j are 1-based as in the reference
// ------------------------------------------------------------------------------------
template <class T>
void canvas_pixel(const MetricSpec& ms, const rtgr_camera& cam, int i, int j, T pos[D], T u[D]) {
    using std::sqrt;
    const T dx = (T(i) - T(1) / 2) / T(cam.ni) - T(1) / 2;
    const T dy = (T(j) - T(1) / 2) / T(cam.nj) - T(1) / 2;
    std::array<T, D> x, n;
    for (int a = 0; a < D; ++a) {
        x[a] = T(cam.pos[a]) + dx * T(cam.widthx[a]) + dy * T(cam.widthy[a]);
        n[a] = T(cam.normal[a]) + dx * T(cam.widthx[a]) + dy * T(cam.widthy[a]);
    }
    Mat4<T> g = eval_metric<T>(x, ms);
    Mat4<T> gu = inverse4<T>(g);
    T t[D];
    for (int a = 0; a < D; ++a) t[a] = gu[a][0];          // gu * (1,0,0,0)
    T t2 = 0, n2 = 0;
    for (int a = 0; a < D; ++a) {
        T gt = 0, gn = 0;
        for (int b = 0; b < D; ++b) { gt += g[a][b] * t[b]; gn += g[a][b] * n[b]; }
        t2 += t[a] * gt; n2 += n[a] * gn;
    }
    const T st = sqrt(-t2), sn = sqrt(n2), s2 = sqrt(T(2));
    for (int a = 0; a < D; ++a) { pos[a] = x[a]; u[a] = (t[a] / st + n[a] / sn) / s2; }
}

// ------------------------------------------------------------------------------------
// Tsit5 (OrdinaryDiffEq 5.38.3 Tsit5ConstantCache)
// ------------------------------------------------------------------------------------
struct Tab {
    static constexpr double a21 = 0.161;
    static constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
    static constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
    static constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                            a54 = -0.09249506636175525;
    static constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159,
                            a64 = -0.071584973281401, a65 = -0.028269050394068383;
    static constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996,
                            a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
    static constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469,
                            bt3 = 0.007880878010261995, bt4 = -0.1447110071732629, bt5 = 0.5823571654525552,
                            bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;
    static constexpr double r11 = 1.0, r12 = -2.763706197274826, r13 = 2.9132554618219126, r14 = -1.0530884977290216;
    static constexpr double r22 = 0.13169999999999998, r23 = -0.2234, r24 = 0.1017;
    static constexpr double r32 = 3.9302962368947516, r33 = -5.941033872131505, r34 = 2.490627285651253;
    static constexpr double r42 = -12.411077166933676, r43 = 30.33818863028232, r44 = -16.548102889244902;
    static constexpr double r52 = 37.50931341651104, r53 = -88.1789048947664, r54 = 47.37952196281928;
    static constexpr double r62 = -27.896526289197286, r63 = 65.09189467479366, r64 = -34.87065786149661;
    static constexpr double r72 = 1.5, r73 = -4.0, r74 = 2.5;
};

template <class T>
State<T> axpy_comb(const State<T>& y0, T dt, std::initializer_list<std::pair<double, const State<T>*>> terms) {
    State<T> out;
    for (int i = 0; i < 2 * D; ++i) {
        T acc = 0; bool first = true;
        for (auto& t : terms) {
            T v = T(t.first) * (*t.second)[i];
            acc = first ? v : acc + v; first = false;
        }
        out[i] = y0[i] + dt * acc;
    }
    return out;
}

template <class T>
void dense_weights(T th, T b[7]) {   // appendix A.6
    const T th2 = th * th;
    b[0] = th * (T(Tab::r11) + th * (T(Tab::r12) + th * (T(Tab::r13) + th * T(Tab::r14))));
    b[1] = th2 * (T(Tab::r22) + th * (T(Tab::r23) + th * T(Tab::r24)));
    b[2] = th2 * (T(Tab::r32) + th * (T(Tab::r33) + th * T(Tab::r34)));
    b[3] = th2 * (T(Tab::r42) + th * (T(Tab::r43) + th * T(Tab::r44)));
    b[4] = th2 * (T(Tab::r52) + th * (T(Tab::r53) + th * T(Tab::r54)));
    b[5] = th2 * (T(Tab::r62) + th * (T(Tab::r63) + th * T(Tab::r64)));
    b[6] = th2 * (T(Tab::r72) + th * (T(Tab::r73) + th * T(Tab::r74)));
}

template <class T>
State<T> dense_eval(T th, T dt, const State<T>& y0, const State<T> k[7]) {
    T b[7]; dense_weights<T>(th, b);
    State<T> out;
    for (int i = 0; i < 2 * D; ++i) {
        T acc = k[0][i] * b[0];
        for (int s = 1; s < 7; ++s) acc = acc + k[s][i] * b[s];
        out[i] = y0[i] + dt * acc;
    }
    return out;
}

template <class T> T rms_norm(const State<T>& v) {   // ODE_DEFAULT_NORM on an array
    using std::sqrt;
    T s = 0; for (int i = 0; i < 2 * D; ++i) s += v[i] * v[i];
    return sqrt(s / T(2 * D));
}

struct RayResult {
    int status;
    int64_t rhs_evals, accepted, rejected;
};

// One trajectory: init dt (A.4), Tsit5 steps (A.1), error norm (A.2), PI controller (A.3),
// event detection + root-find (A.5).  Returns the state the reference would keep as
// sol[end] (src:502-505).
template <class T>
RayResult solve_ray(const MetricSpec& ms, const rtgr_params& P, const rtgr_object* objs, int nobj,
                    const State<T>& u0, State<T>& uend, T* lambda_end) {
    using std::sqrt; using std::fabs; using std::pow; using std::log10; using std::max; using std::min;
    RayResult R{RTGR_STATUS_LAMBDA_END, 0, 0, 0};
    const T abstol = T(P.abstol), reltol = T(P.reltol);
    const T t0 = T(P.lambda0), t1 = T(P.lambda1);
    const T dtmax = t1 - t0;
    const T dtmin = std::numeric_limits<T>::epsilon();
    const T qmin = T(1) / 5, qmax = T(10), gamma = T(9) / 10, beta1 = T(7) / 50, beta2 = T(2) / 25;
    const T qoldinit = T(1e-4);
    auto f = [&](const State<T>& s) { ++R.rhs_evals; return geodesic<T>(s, ms); };
    auto bad = [](const State<T>& s) { for (auto v : s) if (!(v == v)) return true; return false; };

    // ---- initial step size (A.4) ----
    State<T> sk, tmp;
    for (int i = 0; i < 2 * D; ++i) sk[i] = abstol + fabs(u0[i]) * reltol;
    for (int i = 0; i < 2 * D; ++i) tmp[i] = u0[i] / sk[i];
    const T d0 = rms_norm<T>(tmp);
    State<T> f0 = f(u0);
    for (int i = 0; i < 2 * D; ++i) tmp[i] = f0[i] / sk[i];
    const T d1 = rms_norm<T>(tmp);
    T dt0 = (d0 < T(1e-5) || d1 < T(1e-5)) ? T(1e-6) : (d0 / d1) / T(100);
    dt0 = min(dt0, dtmax);
    State<T> u1;
    for (int i = 0; i < 2 * D; ++i) u1[i] = u0[i] + dt0 * f0[i];
    State<T> f1 = f(u1);
    for (int i = 0; i < 2 * D; ++i) tmp[i] = (f1[i] - f0[i]) / sk[i];
    const T d2 = rms_norm<T>(tmp) / dt0;
    const T dm = max(d1, d2);
    T dt1 = (dm <= T(1e-15)) ? max(T(1e-6), dt0 * T(1e-3)) : pow(T(10), -(T(2) + log10(dm)) / T(5));
    T dt = min(min(T(100) * dt0, dt1), dtmax);

    // ---- integration ----
    State<T> uprev = u0, u = u0;
    State<T> k[7];
    k[0] = f(uprev);                       // fsalfirst
    T t = t0, qold = qoldinit, q11 = T(1);
    T cprev = min_distance<T>(objs, nobj, uprev.data());
    int64_t iter = 0;
    while (t < t1) {
        dt = min(dt, t1 - t);              // modify_dt_for_tstops!
        ++iter;
        if (iter > P.maxiters) { R.status = RTGR_STATUS_MAXITERS; break; }
        if (!(fabs(dt) > dtmin)) { R.status = (dt == dt) ? RTGR_STATUS_DT_MIN : RTGR_STATUS_NONFINITE; break; }
        if (bad(uprev)) { R.status = RTGR_STATUS_NONFINITE; break; }
        // stages (A.1)
        k[1] = f(axpy_comb<T>(uprev, dt, {{Tab::a21, &k[0]}}));
        k[2] = f(axpy_comb<T>(uprev, dt, {{Tab::a31, &k[0]}, {Tab::a32, &k[1]}}));
        k[3] = f(axpy_comb<T>(uprev, dt, {{Tab::a41, &k[0]}, {Tab::a42, &k[1]}, {Tab::a43, &k[2]}}));
        k[4] = f(axpy_comb<T>(uprev, dt, {{Tab::a51, &k[0]}, {Tab::a52, &k[1]}, {Tab::a53, &k[2]}, {Tab::a54, &k[3]}}));
        k[5] = f(axpy_comb<T>(uprev, dt, {{Tab::a61, &k[0]}, {Tab::a62, &k[1]}, {Tab::a63, &k[2]}, {Tab::a64, &k[3]},
                                          {Tab::a65, &k[4]}}));
        u = axpy_comb<T>(uprev, dt, {{Tab::a71, &k[0]}, {Tab::a72, &k[1]}, {Tab::a73, &k[2]}, {Tab::a74, &k[3]},
                                     {Tab::a75, &k[4]}, {Tab::a76, &k[5]}});
        k[6] = f(u);
        // embedded error estimate and scaled RMS norm (A.2)
        State<T> ut = axpy_comb<T>(State<T>{}, dt, {{Tab::bt1, &k[0]}, {Tab::bt2, &k[1]}, {Tab::bt3, &k[2]},
                                                    {Tab::bt4, &k[3]}, {Tab::bt5, &k[4]}, {Tab::bt6, &k[5]},
                                                    {Tab::bt7, &k[6]}});
        State<T> res;
        for (int i = 0; i < 2 * D; ++i) res[i] = ut[i] / (abstol + max(fabs(uprev[i]), fabs(u[i])) * reltol);
        const T EEst = rms_norm<T>(res);
        // PI controller (A.3)
        T q;
        if (EEst == 0) {
            q = T(1) / qmax;
        } else {
            q11 = pow(EEst, beta1);
            q = q11 / pow(qold, beta2);
            q = max(T(1) / qmax, min(T(1) / qmin, q / gamma));
        }
        if (!(EEst <= T(1))) {             // reject (NaN EEst lands here too)
            ++R.rejected;
            dt = dt / min(T(1) / qmin, q11 / gamma);
            continue;
        }
        ++R.accepted;
        // qsteady_min = qsteady_max = 1: the "q -> 1 inside the window" rule is a no-op
        qold = max(EEst, qoldinit);
        const T dtnew = min(dt / q, dtmax);
        const T tprev = t;
        const T ttmp = t + dt;
        t = (fabs(ttmp - t1) < T(10) * std::numeric_limits<T>::epsilon() * max(ttmp, t1)) ? t1 : ttmp;

        // ---- ContinuousCallback (A.5) ----
        const T c0 = cprev;
        const T c1 = min_distance<T>(objs, nobj, u.data());
        const int np = P.interp_points;
        bool event = false; T th_lo = 0, th_hi = 1;
        const T s0 = (c0 > 0) ? T(1) : (c0 < 0 ? T(-1) : T(0));
        const T s1 = (c1 > 0) ? T(1) : (c1 < 0 ? T(-1) : T(0));
        if (s0 != 0 && s0 * s1 <= 0) {
            event = true;
        } else if (np > 1 && s0 != 0) {
            T prev_th = 0;
            for (int i = 1; i <= np - 2; ++i) {          // interior sample points
                const T th = T(i) / T(np - 1);
                State<T> ui = dense_eval<T>(th, dt, uprev, k);
                const T ci = min_distance<T>(objs, nobj, ui.data());
                if (s0 * ci < 0) { event = true; th_lo = prev_th; th_hi = th; break; }
                prev_th = th;
            }
        }
        if (event) {
            // bracketing root-find on theta -> cond(interp(theta)), driven to bracket collapse; the
            // reference takes the float just before the crossing (prevfloat of the Roots.jl result).
            auto cond_at = [&](T th) {
                if (th == T(1)) return c1;
                if (th == T(0)) return c0;
                State<T> ui = dense_eval<T>(th, dt, uprev, k);
                return min_distance<T>(objs, nobj, ui.data());
            };
            T lo = th_lo, hi = th_hi;
            T chi = cond_at(hi);
            T th_star;
            if (chi == 0) {
                th_star = hi;
            } else {
                T clo = cond_at(lo);
                for (int it = 0; it < 200; ++it) {
                    T mid;
                    // secant proposal safeguarded by bisection
                    T sec = lo - clo * (hi - lo) / (chi - clo);
                    if ((it & 1) == 0 && sec > lo && sec < hi) mid = sec; else mid = lo + (hi - lo) / 2;
                    if (!(mid > lo && mid < hi)) break;     // adjacent floats: collapsed
                    T cm = cond_at(mid);
                    if (cm == 0) { lo = mid; clo = cm; break; }
                    if (s0 * cm > 0) { lo = mid; clo = cm; } else { hi = mid; chi = cm; }
                }
                th_star = lo;
            }
            if (th_star == T(1)) uend = u;
            else if (th_star == T(0)) uend = uprev;
            else uend = dense_eval<T>(th_star, dt, uprev, k);
            t = tprev + th_star * dt;
            R.status = RTGR_STATUS_EVENT;
            if (lambda_end) *lambda_end = t;
            return R;
        }
        // accept: FSAL
        uprev = u; k[0] = k[6]; cprev = c1;
        dt = dtnew;
    }
    uend = uprev;
    if (R.status == RTGR_STATUS_LAMBDA_END && !(t >= t1)) R.status = RTGR_STATUS_NONFINITE;
    if (lambda_end) *lambda_end = t;
    return R;
}

// classification + colouring (src:513-533)
template <class T>
int classify(const rtgr_params& P, const rtgr_object* objs, int nobj, const T* x, T col[3]) {
    int omin = 0;
    T dmin = T(P.hit_threshold);                 // src:519
    for (int o = 0; o < nobj; ++o) {
        T d = obj_distance<T>(objs[o], x);
        if (d < dmin) { omin = o + 1; dmin = d; }   // src:522-525
    }
    if (omin == 0) { col[0] = 1; col[1] = 0; col[2] = 0; }   // src:528
    else {
        obj_color<T>(objs[omin - 1], x, col);
        const T w = T(omin) / T(nobj);               // src:530
        for (int c = 0; c < 3; ++c) col[c] *= w;
    }
    return omin;
}

MetricSpec spec_of(const rtgr_params& p) { return MetricSpec{p.metric, p.r_formula, p.M, p.a}; }

template <class T>
int trace_impl(const rtgr_params* P, const rtgr_object* objs, int nobj, double* pixels, int64_t n,
               double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps, rtgr_stats* stats,
               int nthreads) {
    const MetricSpec ms = spec_of(*P);
    int64_t rhs = 0, acc = 0, rej = 0;
    auto w0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    // static contiguous chunks over the linear pixel index: what DiffEqBase's EnsembleThreads
    // (the default ensemble algorithm behind src:510) does with Threads.@threads
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(+ : rhs, acc, rej)
    for (int64_t i = 0; i < n; ++i) {
        double* px = pixels + 11 * i;
        State<T> u0, ue;
        for (int a = 0; a < 8; ++a) u0[a] = T(px[a]);                  // src:492-496
        RayResult r = solve_ray<T>(ms, *P, objs, nobj, u0, ue, nullptr);
        T col[3];
        int om = classify<T>(*P, objs, nobj, ue.data(), col);
        px[8] = double(col[0]); px[9] = double(col[1]); px[10] = double(col[2]);   // src:532
        if (final_state) for (int a = 0; a < 8; ++a) final_state[8 * i + a] = double(ue[a]);
        if (obj_id) obj_id[i] = om;
        if (status) status[i] = r.status;
        if (nsteps) nsteps[i] = int32_t(r.accepted);
        rhs += r.rhs_evals; acc += r.accepted; rej += r.rejected;
    }
    auto w1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->rays = uint64_t(n); stats->rhs_evals = uint64_t(rhs);
        stats->steps_accepted = uint64_t(acc); stats->steps_rejected = uint64_t(rej);
        stats->kernel_ms = std::chrono::duration<double, std::milli>(w1 - w0).count();
        stats->total_ms = stats->kernel_ms;
        stats->drain_ms = 0.0;
    }
    return 0;
}

}  // namespace

// =====================================================================================
// C entry points (ctypes-facing; tests and the bench's CPU-baseline leg only)
// =====================================================================================
extern "C" {

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// g: 16 doubles row-major
int oracle_metric(const rtgr_params* P, const double* x, double* g) {
    std::array<double, D> xx{x[0], x[1], x[2], x[3]};
    Mat4<double> m = eval_metric<double>(xx, spec_of(*P));
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) g[4 * a + b] = m[a][b];
    return 0;
}

// g[a][b] row-major 16, dg[a][b][c] 64 (c fastest)
int oracle_dmetric(const rtgr_params* P, const double* x, double* g, double* dg) {
    std::array<double, D> xx{x[0], x[1], x[2], x[3]};
    Mat4<double> m; double d[D][D][D]; bool ok;
    dmetric<double>(spec_of(*P), xx, m, d, &ok);
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) {
        g[4 * a + b] = m[a][b];
        for (int c = 0; c < D; ++c) dg[16 * a + 4 * b + c] = d[a][b][c];
    }
    return ok ? 0 : 1;
}

int oracle_christoffel(const rtgr_params* P, const double* x, double* Gam) {
    std::array<double, D> xx{x[0], x[1], x[2], x[3]};
    double G[D][D][D]; bool ok;
    christoffel<double>(spec_of(*P), xx, G, &ok);
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) for (int c = 0; c < D; ++c)
        Gam[16 * a + 4 * b + c] = G[a][b][c];
    return ok ? 0 : 1;
}

int oracle_inverse4(const double* A, double* B, double* det) {
    Mat4<double> m;
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) m[a][b] = A[4 * a + b];
    Mat4<double> r = inverse4<double>(m, det);
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) B[4 * a + b] = r[a][b];
    return 0;
}

// The Float32 checks of the reference's "Kerr-Schild metric" testset (test/runtests.jl:36-61):
// out = { any NaN in g, |det g * det g^-1 - 1|, max|g g^-1 - I|, max|dmetric.g - metric|, any NaN in Gamma }
int oracle_ks_checks_f32(const rtgr_params* P, const float* x, float* out) {
    MetricSpec ms = spec_of(*P);
    std::array<float, D> xx{x[0], x[1], x[2], x[3]};
    Mat4<float> g = eval_metric<float>(xx, ms);
    float nan_g = 0; for (auto& r : g) for (auto v : r) if (v != v) nan_g = 1;
    float detg, detgu;
    Mat4<float> gu = inverse4<float>(g, &detg);
    inverse4<float>(gu, &detgu);
    float maxdev = 0;
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) {
        float s = 0; for (int c = 0; c < D; ++c) s += g[a][c] * gu[c][b];
        maxdev = std::max(maxdev, std::fabs(s - (a == b ? 1.f : 0.f)));
    }
    Mat4<float> g1; float dg[D][D][D]; bool ok;
    dmetric<float>(ms, xx, g1, dg, &ok);
    float gdev = 0;
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) gdev = std::max(gdev, std::fabs(g1[a][b] - g[a][b]));
    float G[D][D][D]; christoffel<float>(ms, xx, G, &ok);
    float nan_G = 0;
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) for (int c = 0; c < D; ++c) if (G[a][b][c] != G[a][b][c]) nan_G = 1;
    out[0] = nan_g; out[1] = std::fabs(detg * detgu - 1.f); out[2] = maxdev; out[3] = gdev; out[4] = nan_G;
    return 0;
}

int oracle_rhs_batch(const rtgr_params* P, const double* states, int64_t n, double* derivs) {
    const MetricSpec ms = spec_of(*P);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        State<double> s;
        for (int a = 0; a < 8; ++a) s[a] = states[8 * i + a];
        State<double> d = geodesic<double>(s, ms);
        for (int a = 0; a < 8; ++a) derivs[8 * i + a] = d[a];
    }
    return 0;
}

// same in extended precision (x87 80-bit), returned rounded to double: a truth estimate
int oracle_rhs_batch_ld(const rtgr_params* P, const double* states, int64_t n, double* derivs) {
    const MetricSpec ms = spec_of(*P);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        State<long double> s;
        for (int a = 0; a < 8; ++a) s[a] = states[8 * i + a];
        State<long double> d = geodesic<long double>(s, ms);
        for (int a = 0; a < 8; ++a) derivs[8 * i + a] = double(d[a]);
    }
    return 0;
}

int oracle_make_canvas(const rtgr_params* P, const rtgr_camera* cam, double* pixels) {
    const MetricSpec ms = spec_of(*P);
    const int ni = cam->ni, nj = cam->nj;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= nj; ++j)
        for (int i = 1; i <= ni; ++i) {
            double* px = pixels + 11 * (int64_t(i - 1) + int64_t(j - 1) * ni);
            canvas_pixel<double>(ms, *cam, i, j, px, px + 4);
            px[8] = px[9] = px[10] = 0.0;
        }
    return 0;
}

int oracle_trace_pixels(const rtgr_params* P, const rtgr_object* objs, int nobj, double* pixels, int64_t n,
                        double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                        rtgr_stats* stats, int nthreads) {
    return trace_impl<double>(P, objs, nobj, pixels, n, final_state, obj_id, status, nsteps, stats, nthreads);
}

// extended-precision run of the same algorithm (truth estimate for the parity analysis)
int oracle_trace_pixels_ld(const rtgr_params* P, const rtgr_object* objs, int nobj, double* pixels, int64_t n,
                           double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                           rtgr_stats* stats, int nthreads) {
    return trace_impl<long double>(P, objs, nobj, pixels, n, final_state, obj_id, status, nsteps, stats, nthreads);
}

// rgb8: nj x ni x 3 row-major (row = j, col = i), value = round(255 x) -- the PNG the
// reference saves (src:566-575)
int oracle_quantize(const double* pixels, int ni, int nj, uint8_t* rgb8) {
    for (int j = 0; j < nj; ++j)
        for (int i = 0; i < ni; ++i)
            for (int c = 0; c < 3; ++c) {
                double v = pixels[11 * (int64_t(i) + int64_t(j) * ni) + 8 + c];
                v = std::min(1.0, std::max(0.0, v));
                rgb8[(int64_t(j) * ni + i) * 3 + c] = uint8_t(std::nearbyint(255.0 * v));
            }
    return 0;
}

}  // extern "C"
