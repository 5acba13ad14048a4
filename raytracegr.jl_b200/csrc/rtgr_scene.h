// Host-side flattening of the ABI structs (params, objects, camera) into the SceneConst block
// the kernels read from __constant__ memory.  Also validates the arguments.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "rtgr_core.cuh"

namespace rtgr {

inline bool build_scene_const(const rtgr_params* p, const rtgr_object* objs, int n_objs,
                              const rtgr_camera* cam, SceneConst& sc, std::string& err) {
    std::memset(&sc, 0, sizeof(sc));
    if (!p) { err = "params is NULL"; return false; }
    if (p->metric != RTGR_MINKOWSKI && p->metric != RTGR_KERR_SCHILD && p->metric < RTGR_USER_METRIC_BASE) {
        err = "unknown metric kind"; return false;
    }
    if (p->r_formula != RTGR_R_AS_WRITTEN && p->r_formula != RTGR_R_CORRECTED) { err = "unknown r_formula"; return false; }
    if (n_objs < 0 || n_objs > RTGR_MAX_OBJECTS) { err = "n_objs out of range (0..16)"; return false; }
    if (n_objs > 0 && !objs) { err = "objs is NULL"; return false; }
    if (p->interp_points < 0 || p->interp_points > MAX_INTERP) { err = "interp_points out of range (0..32)"; return false; }
    if (!(p->reltol > 0.0) || !(p->abstol > 0.0)) { err = "tolerances must be positive"; return false; }
    if (!(p->lambda1 > p->lambda0)) { err = "lambda1 must exceed lambda0"; return false; }
    sc.M = p->M; sc.a = p->a; sc.a2 = p->a * p->a; sc.twoM = 2.0 * p->M; sc.twoa2 = 2.0 * p->a * p->a;
    sc.lambda0 = p->lambda0; sc.lambda1 = p->lambda1;
    sc.reltol = p->reltol; sc.abstol = p->abstol; sc.hit_threshold = p->hit_threshold;
    sc.dtmax = p->lambda1 - p->lambda0;
    {   // see step_far_from_end: signed high-word order is the order of the values only for positive numbers
        const double h = 0.5 * p->lambda1, q = 0.25 * p->lambda1;
        int64_t bh, bq;
        std::memcpy(&bh, &h, 8); std::memcpy(&bq, &q, 8);
        const bool ok = p->lambda1 > 0.0 && p->lambda0 >= 0.0;
        sc.t1_half_hi = ok ? int32_t(bh >> 32) : INT32_MIN;
        sc.t1_quarter_hi = ok ? int32_t(bq >> 32) : INT32_MIN;
    }
    sc.interp_points = p->interp_points; sc.maxiters = p->maxiters; sc.n_objs = n_objs; sc.metric = p->metric;
    for (int i = 0; i < p->interp_points; ++i) sc.theta[i] = double(i) / double(p->interp_points - 1);
    for (int o = 0; o < n_objs; ++o) {
        const rtgr_object& ob = objs[o];
        if (ob.kind != RTGR_PLANE && ob.kind != RTGR_SPHERE) { err = "unknown object kind"; return false; }
        sc.kind[o] = ob.kind;
        sc.time[o] = ob.time;
        sc.cx[o] = ob.pos[1]; sc.cy[o] = ob.pos[2]; sc.cz[o] = ob.pos[3];
        sc.R2[o] = ob.radius * ob.radius;
        sc.Rabs[o] = std::fabs(ob.radius);
        if (ob.kind == RTGR_PLANE) {
            sc.qa[o] = 0.0; sc.qb0[o] = 1.0; sc.qb1[o] = sc.qb2[o] = sc.qb3[o] = 0.0; sc.qc[o] = -ob.time;
            sc.mA[o] = 1.0; sc.mB[o] = 0.0;
        } else {
            const double sg = ob.radius > 0 ? 1.0 : (ob.radius < 0 ? -1.0 : 0.0);
            sc.qa[o] = sg; sc.qb0[o] = 0.0;
            sc.qb1[o] = -2.0 * sg * ob.pos[1]; sc.qb2[o] = -2.0 * sg * ob.pos[2]; sc.qb3[o] = -2.0 * sg * ob.pos[3];
            sc.qc[o] = sg * (ob.pos[1] * ob.pos[1] + ob.pos[2] * ob.pos[2] + ob.pos[3] * ob.pos[3] - sc.R2[o]);
            sc.mA[o] = 2.0 * std::fabs(ob.radius) * 1.7320508075688774; sc.mB[o] = 3.0;
        }
        sc.sgn[o] = ob.radius > 0 ? 1.0 : (ob.radius < 0 ? -1.0 : 0.0);
    }
    for (int o = n_objs; o < RTGR_MAX_OBJECTS; ++o) sc.qc[o] = INFINITY;   // padding: distance +inf (min_distance_q4)
    for (int o = 0; o < n_objs; ++o) {
        sc.qa_pos_max = std::fmax(sc.qa_pos_max, sc.qa[o]);
        sc.mA_max = std::fmax(sc.mA_max, sc.mA[o]);
        sc.mB_max = std::fmax(sc.mB_max, sc.mB[o]);
    }
    sc.qa_pos_max_q = 0.25 * sc.qa_pos_max;
    sc.nobj_d = double(n_objs);
    sc.inv_nobj = n_objs ? 1.0 / n_objs : 0.0;
    if (cam) {
        if (cam->ni <= 0 || cam->nj <= 0) { err = "camera ni/nj must be positive"; return false; }
        for (int c = 0; c < 4; ++c) {
            sc.cam_pos[c] = cam->pos[c]; sc.cam_wx[c] = cam->widthx[c];
            sc.cam_wy[c] = cam->widthy[c]; sc.cam_n[c] = cam->normal[c];
        }
        sc.ni = cam->ni; sc.nj = cam->nj;
    }
    return true;
}

// Tile bookkeeping for render jobs: tiles t = offset + m*stride, m = 0..count-1.
inline void tile_selection(int ni, int nj, int offset, int stride, int& tiles_x, int64_t& count) {
    tiles_x = (ni + RTGR_TILE_W - 1) / RTGR_TILE_W;
    const int tiles_y = (nj + RTGR_TILE_H - 1) / RTGR_TILE_H;
    const int64_t ntiles = int64_t(tiles_x) * tiles_y;
    count = (offset < ntiles) ? (ntiles - offset + stride - 1) / stride : 0;
}

// Queue order for render jobs in a Kerr-Schild scene: tiles sorted by the impact parameter of their
// centre ray with respect to the hole (which sits at the spatial origin of Kerr-Schild coordinates),
// smallest first.  Rays that pass close to the hole -- captured or nearly critical -- take 10-30x
// more steps than far-field rays, so handing them out first leaves only cheap, uniform rays for the
// end of the launch (longest-processing-time-first): the drain tail shrinks from the duration of the
// longest ray to that of the shortest.  Interleaving the sorted list over ranks/devices also gives
// every shard the same cost mix.  Purely a schedule: results do not depend on it.
inline std::vector<int32_t> tiles_sorted_by_key(const std::vector<double>& key) {
    std::vector<int32_t> order(key.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    return order;
}

inline std::vector<double> tile_impact_keys(const rtgr_camera& cam) {
    const int tiles_x = (cam.ni + RTGR_TILE_W - 1) / RTGR_TILE_W;
    const int tiles_y = (cam.nj + RTGR_TILE_H - 1) / RTGR_TILE_H;
    const int n = tiles_x * tiles_y;
    std::vector<double> key(n);
    for (int t = 0; t < n; ++t) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const double ci = std::min(double(cam.ni), tx * double(RTGR_TILE_W) + 0.5 * RTGR_TILE_W);
        const double cj = std::min(double(cam.nj), ty * double(RTGR_TILE_H) + 0.5 * RTGR_TILE_H);
        const double dx = ci / cam.ni - 0.5, dy = cj / cam.nj - 0.5;
        double x[3], d[3];
        for (int c = 0; c < 3; ++c) {
            x[c] = cam.pos[c + 1] + dx * cam.widthx[c + 1] + dy * cam.widthy[c + 1];
            d[c] = cam.normal[c + 1] + dx * cam.widthx[c + 1] + dy * cam.widthy[c + 1];
        }
        const double xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        const double xd = x[0] * d[0] + x[1] * d[1] + x[2] * d[2];
        const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        // squared distance of closest approach of the straight half-line the ray starts on
        const double k = (dd > 0.0 && xd < 0.0) ? xx - xd * xd / dd : xx;
        key[t] = (k == k) ? k : 0.0;   // NaN camera: a key that still orders (std::stable_sort needs a strict weak order)
    }
    return key;
}
inline std::vector<int32_t> tile_order_by_impact(const rtgr_camera& cam) { return tiles_sorted_by_key(tile_impact_keys(cam)); }

// The same order when the rays come from a caller-supplied Pixel array (rtgr_trace_canvas): the key is
// taken from the pixel nearest each tile's centre (pos = start point, normal = initial 4-velocity,
// whose spatial part is the ray direction).
inline std::vector<double> tile_impact_keys_pixels(const rtgr_pixel* px, int ni, int nj) {
    const int tiles_x = (ni + RTGR_TILE_W - 1) / RTGR_TILE_W;
    const int tiles_y = (nj + RTGR_TILE_H - 1) / RTGR_TILE_H;
    const int n = tiles_x * tiles_y;
    std::vector<double> key(n);
    for (int t = 0; t < n; ++t) {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int ci = std::min(ni - 1, tx * RTGR_TILE_W + RTGR_TILE_W / 2);
        const int cj = std::min(nj - 1, ty * RTGR_TILE_H + RTGR_TILE_H / 2);
        const rtgr_pixel& q = px[int64_t(ci) + int64_t(cj) * ni];
        const double* x = q.pos + 1;
        const double* d = q.normal + 1;
        const double xx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        const double xd = x[0] * d[0] + x[1] * d[1] + x[2] * d[2];
        const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double k = (dd > 0.0 && xd < 0.0) ? xx - xd * xd / dd : xx;
        key[t] = (k == k) ? k : 0.0;   // NaN input: treat as expensive
    }
    return key;
}
inline std::vector<int32_t> tile_order_by_impact_pixels(const rtgr_pixel* px, int ni, int nj) {
    return tiles_sorted_by_key(tile_impact_keys_pixels(px, ni, nj));
}

}  // namespace rtgr
