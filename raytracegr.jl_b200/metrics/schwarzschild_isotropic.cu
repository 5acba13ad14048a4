// Schwarzschild black hole in isotropic coordinates (a metric the reference does not ship; any
// callable is accepted by its trace_rays, src:483):
//   ds^2 = -((1 - m/2rho)/(1 + m/2rho))^2 dt^2 + (1 + m/2rho)^4 (dx^2 + dy^2 + dz^2),  par[0] = m.
template <class T>
__device__ void rtgr_user_metric(const T xx[4], T g[4][4], const double* par) {
    const double m = par[0];
    const T rho = sqrt(pow2(xx[1]) + pow2(xx[2]) + pow2(xx[3]));
    const T w = m / (2 * rho);
    const T psi = 1 + w;
    for (int p = 0; p < 4; ++p)
        for (int q = 0; q < 4; ++q) g[p][q] = T(0);
    g[0][0] = -pow2((1 - w) / psi);
    g[1][1] = g[2][2] = g[3][3] = pow4(psi);
}
