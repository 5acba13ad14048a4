// The reference's kerr_schild (src/RayTraceGR.jl:274-294) handed to rtgr_metric_compile in KERR-SCHILD FORM: the
// author supplies the scalar f and the covector k of  g = eta + f k (x) k  (src:285-291) instead of the 4x4 matrix,
// and declares that nothing depends on x[0].  The library then evaluates the geodesic right-hand side in closed
// form (no 4x4 derivative sets, no matrix inverse).  Same formulas as kerr_schild_as_written.cu, line for line.
// par[0] = M, par[1] = a  (the reference hard-codes 1 and 0, src:275-276).
#pragma rtgr stationary
template <class T>
__device__ void rtgr_user_kerr_schild(const T xx[4], T& f, T k[4], const double* par) {
    const double M = par[0], a = par[1];
    const T x = xx[1], y = xx[2], z = xx[3];
    const T rho = sqrt(pow2(x) + pow2(y) + pow2(z));                                     // src:283
    const T r = sqrt(pow2(rho) - a * a) / 2 + sqrt(a * a * pow2(z) + pow2((pow2(rho) - a * a) / 2));   // src:284
    f = 2 * M * pow3(r) / (pow4(r) + a * a * pow2(z));                                   // src:285
    k[0] = T(1);                                                                          // src:286
    k[1] = (r * x + a * y) / (pow2(r) + a * a);                                           // src:287
    k[2] = (r * y - a * x) / (pow2(r) + a * a);                                           // src:288
    k[3] = z / r;                                                                         // src:289
}
