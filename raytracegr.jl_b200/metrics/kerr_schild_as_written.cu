// The reference's kerr_schild (src/RayTraceGR.jl:274-294) written as a USER metric for
// rtgr_metric_compile: literally the reference's lines, including the radius at src:284.
// par[0] = M, par[1] = a  (the reference hard-codes 1 and 0, src:275-276).
template <class T>
__device__ void rtgr_user_metric(const T xx[4], T g[4][4], const double* par) {
    const double M = par[0], a = par[1];
    const T x = xx[1], y = xx[2], z = xx[3];
    const T rho = sqrt(pow2(x) + pow2(y) + pow2(z));                                     // src:283
    const T r = sqrt(pow2(rho) - a * a) / 2 + sqrt(a * a * pow2(z) + pow2((pow2(rho) - a * a) / 2));   // src:284
    const T f = 2 * M * pow3(r) / (pow4(r) + a * a * pow2(z));                           // src:285
    T k[4];
    k[0] = T(1);                                                                          // src:286
    k[1] = (r * x + a * y) / (pow2(r) + a * a);                                           // src:287
    k[2] = (r * y - a * x) / (pow2(r) + a * a);                                           // src:288
    k[3] = z / r;                                                                         // src:289
    for (int p = 0; p < 4; ++p)
        for (int q = 0; q < 4; ++q) {
            const double eta = (p == q) ? (p == 0 ? -1.0 : 1.0) : 0.0;                   // src:262-264
            g[p][q] = eta + f * k[p] * k[q];                                              // src:291
        }
}
