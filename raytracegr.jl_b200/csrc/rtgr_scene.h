// Host-side flattening of the ABI structs (params, objects, camera) into the SceneConst block
// the kernels read from __constant__ memory.  Also validates the arguments.
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "rtgr_core.cuh"

namespace rtgr {

inline bool build_scene_const(const rtgr_params* p, const rtgr_object* objs, int n_objs,
                              const rtgr_camera* cam, SceneConst& sc, std::string& err) {
    std::memset(&sc, 0, sizeof(sc));
    if (!p) { err = "params is NULL"; return false; }
    if (p->metric != RTGR_MINKOWSKI && p->metric != RTGR_KERR_SCHILD) { err = "unknown metric kind"; return false; }
    if (p->r_formula != RTGR_R_AS_WRITTEN && p->r_formula != RTGR_R_CORRECTED) { err = "unknown r_formula"; return false; }
    if (n_objs < 0 || n_objs > RTGR_MAX_OBJECTS) { err = "n_objs out of range (0..16)"; return false; }
    if (n_objs > 0 && !objs) { err = "objs is NULL"; return false; }
    if (p->interp_points < 0 || p->interp_points > MAX_INTERP) { err = "interp_points out of range (0..32)"; return false; }
    if (!(p->reltol > 0.0) || !(p->abstol > 0.0)) { err = "tolerances must be positive"; return false; }
    if (!(p->lambda1 > p->lambda0)) { err = "lambda1 must exceed lambda0"; return false; }
    sc.M = p->M; sc.a = p->a; sc.a2 = p->a * p->a; sc.twoM = 2.0 * p->M;
    sc.lambda0 = p->lambda0; sc.lambda1 = p->lambda1;
    sc.reltol = p->reltol; sc.abstol = p->abstol; sc.hit_threshold = p->hit_threshold;
    sc.dtmax = p->lambda1 - p->lambda0;
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
    for (int o = 0; o < n_objs; ++o) {
        sc.qa_pos_max = std::fmax(sc.qa_pos_max, sc.qa[o]);
        sc.mA_max = std::fmax(sc.mA_max, sc.mA[o]);
        sc.mB_max = std::fmax(sc.mB_max, sc.mB[o]);
    }
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

}  // namespace rtgr
