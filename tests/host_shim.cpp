// TEST SCAFFOLDING: compiles the product's per-lane arithmetic (csrc/rtgr_core.cuh,
// csrc/rtgr_trace.cuh) for the host with a trivial one-lane scheduler, so the integrator logic can
// be unit-tested against the oracle on machines without a GPU.  Never shipped, never used by the
// product path (which has no CPU fallback).
#include <atomic>
#include <string>
#include <vector>

#include "../raytracegr.jl_b200/csrc/rtgr_scene.h"
#include "../raytracegr.jl_b200/csrc/rtgr_trace.cuh"

namespace {
struct HostSched {
    int64_t* next;
    bool any(bool p) const { return p; }
    bool all(bool p) const { return p; }
    int64_t fetch(bool want, const rtgr::Job&) { return want ? (*next)++ : -1; }
    static constexpr bool STAGE = false;     // (the RGB8 patch staging and the chunk-wise ray reads are the CUDA scheduler's)
    static constexpr bool PREFETCH = false;
    void take_rays(const rtgr::Job&, bool, int64_t, int, double*) {}
    static int put_rgbf(const rtgr::SceneConst&, const rtgr::Job&, int32_t, const double*) { return -1; }
    void stage_refill(const rtgr::Job&, int) {}
    static int put_rgb8(const rtgr::SceneConst&, const rtgr::Job&, int32_t, uint32_t) { return -1; }
};

struct HostAcc {   // a VIEW of the lane's seven stage accelerations (copies share the storage, like csrc's SmemAcc)
    using Backing = HostAcc;
    HostAcc backing() const { return *this; }
    double (*A)[4];
    void load(int i, double v[4]) const { for (int c = 0; c < 4; ++c) v[c] = A[i][c]; }
    void store(int i, const double v[4]) const { for (int c = 0; c < 4; ++c) A[i][c] = v[c]; }
};

const rtgr::StageTab g_tab = rtgr::make_stage_tab();

// One lane drawing 32-ordinal chunks (one 8x4-pixel patch) from a queue head that OTHER PROCESSES draw
// from as well -- the head lives in shared memory.  Host-side model of the cross-GPU tile queue of
// rtgr_render_frame (csrc WarpSched with Job::queue_scope = 1), for the world-size-2 CPU test.
struct SharedQueueSched {
    unsigned long long* head;
    int64_t c_base = 0;
    int c_left = 0;
    bool any(bool p) const { return p; }
    bool all(bool p) const { return p; }
    int64_t fetch(bool want, const rtgr::Job&) {
        if (!want) return -1;
        if (c_left == 0) { c_base = int64_t(__atomic_fetch_add(head, 32ull, __ATOMIC_RELAXED)); c_left = 32; }
        --c_left;
        return c_base++;
    }
    static constexpr bool STAGE = false;     // (the RGB8 patch staging and the chunk-wise ray reads are the CUDA scheduler's)
    static constexpr bool PREFETCH = false;
    void take_rays(const rtgr::Job&, bool, int64_t, int, double*) {}
    static int put_rgbf(const rtgr::SceneConst&, const rtgr::Job&, int32_t, const double*) { return -1; }
    void stage_refill(const rtgr::Job&, int) {}
    static int put_rgb8(const rtgr::SceneConst&, const rtgr::Job&, int32_t, uint32_t) { return -1; }
};

template <int METRIC, int RFORM, class Sched>
void run(const rtgr::SceneConst& sc, const rtgr::Job& job, rtgr::Counters& cnt, Sched& s) {
    double storage[7][4];
    HostAcc acc{storage};
    rtgr::trace_loop<METRIC, RFORM, Sched, HostAcc>(sc, g_tab, job, s, acc, cnt);
}

template <class Sched>
void dispatch_with(const rtgr::SceneConst& sc, int rform, const rtgr::Job& job, rtgr::Counters& cnt, Sched& s) {
    // same selection as csrc's variant_of: a == 0 runs the specialised right-hand side
    if (sc.metric == RTGR_MINKOWSKI) run<RTGR_MINKOWSKI, RTGR_R_AS_WRITTEN>(sc, job, cnt, s);
    else if (sc.a == 0.0 && rform == RTGR_R_AS_WRITTEN) run<RTGR_KERR_SCHILD, rtgr::RFORM_A0 + RTGR_R_AS_WRITTEN>(sc, job, cnt, s);
    else if (sc.a == 0.0) run<RTGR_KERR_SCHILD, rtgr::RFORM_A0 + RTGR_R_CORRECTED>(sc, job, cnt, s);
    else if (rform == RTGR_R_AS_WRITTEN) run<RTGR_KERR_SCHILD, RTGR_R_AS_WRITTEN>(sc, job, cnt, s);
    else run<RTGR_KERR_SCHILD, RTGR_R_CORRECTED>(sc, job, cnt, s);
}

void dispatch(const rtgr::SceneConst& sc, int rform, const rtgr::Job& job, rtgr::Counters& cnt) {
    int64_t next = 0;
    HostSched s{&next};
    dispatch_with(sc, rform, job, cnt, s);
}
}  // namespace

extern "C" {

int shim_rhs_batch(const rtgr_params* p, const double* states, int64_t n, double* derivs) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, nullptr, 0, nullptr, sc, err)) return -1;
    for (int64_t i = 0; i < n; ++i) {
        double A[4];
        if (p->metric == RTGR_MINKOWSKI) rtgr::accel<RTGR_MINKOWSKI, 0>(sc, states + 8 * i, A);
        else if (p->a == 0.0 && p->r_formula == RTGR_R_AS_WRITTEN) rtgr::accel<RTGR_KERR_SCHILD, rtgr::RFORM_A0 + RTGR_R_AS_WRITTEN>(sc, states + 8 * i, A);
        else if (p->a == 0.0) rtgr::accel<RTGR_KERR_SCHILD, rtgr::RFORM_A0 + RTGR_R_CORRECTED>(sc, states + 8 * i, A);
        else if (p->r_formula == RTGR_R_AS_WRITTEN) rtgr::accel<RTGR_KERR_SCHILD, RTGR_R_AS_WRITTEN>(sc, states + 8 * i, A);
        else rtgr::accel<RTGR_KERR_SCHILD, RTGR_R_CORRECTED>(sc, states + 8 * i, A);
        for (int c = 0; c < 4; ++c) { derivs[8 * i + c] = states[8 * i + 4 + c]; derivs[8 * i + 4 + c] = A[c]; }
    }
    return 0;
}

int shim_make_canvas(const rtgr_params* p, const rtgr_camera* cam, double* pixels) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, nullptr, 0, cam, sc, err)) return -1;
    for (int j = 0; j < cam->nj; ++j)
        for (int i = 0; i < cam->ni; ++i) {
            double* px = pixels + 11 * (int64_t(i) + int64_t(j) * cam->ni);
            if (p->metric == RTGR_MINKOWSKI) rtgr::canvas_pixel<RTGR_MINKOWSKI, 0>(sc, i, j, px, px + 4);
            else if (p->r_formula == RTGR_R_AS_WRITTEN) rtgr::canvas_pixel<RTGR_KERR_SCHILD, RTGR_R_AS_WRITTEN>(sc, i, j, px, px + 4);
            else rtgr::canvas_pixel<RTGR_KERR_SCHILD, RTGR_R_CORRECTED>(sc, i, j, px, px + 4);
            px[8] = px[9] = px[10] = 0.0;
        }
    return 0;
}

int shim_trace_pixels(const rtgr_params* p, const rtgr_object* objs, int n_objs, const double* pixels, int64_t n,
                      double* rgb_f64, double* final_state, int32_t* obj_id, int32_t* status, int32_t* nsteps,
                      uint64_t* counters /* rays, attempts, accepted, rejected */) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, objs, n_objs, nullptr, sc, err)) return -1;
    rtgr::Job job{};
    job.mode = rtgr::JOB_PIXELS; job.total = n; job.pixels_in = pixels;
    job.rgb_f64 = rgb_f64; job.rgb_stride = 3; job.final_state = final_state; job.obj_id = obj_id; job.status = status; job.nsteps = nsteps;
    rtgr::Counters cnt{0, 0, 0, 0};
    dispatch(sc, p->r_formula, job, cnt);
    if (counters) { counters[0] = cnt.rays; counters[1] = cnt.attempts; counters[2] = cnt.accepted; counters[3] = cnt.rejected; }
    return 0;
}

int shim_render_tiles(const rtgr_params* p, const rtgr_object* objs, int n_objs, const rtgr_camera* cam,
                      int tile_offset, int tile_stride, uint8_t* rgb8, double* rgb_f64, double* final_state,
                      int32_t* obj_id, int32_t* status, int32_t* nsteps, uint64_t* counters) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, objs, n_objs, cam, sc, err)) return -1;
    rtgr::Job job{};
    job.mode = rtgr::JOB_RENDER; job.tile_offset = tile_offset; job.tile_stride = tile_stride;
    int64_t count;
    rtgr::tile_selection(cam->ni, cam->nj, tile_offset, tile_stride, job.tiles_x, count);
    job.total = count * (RTGR_TILE_W * RTGR_TILE_H);
    std::vector<int32_t> order;
    if (p->metric == RTGR_KERR_SCHILD) { order = rtgr::tile_order_by_impact(*cam); job.tile_order = order.data(); }
    job.rgb8 = rgb8; job.rgb_f64 = rgb_f64; job.rgb_stride = 3; job.final_state = final_state; job.obj_id = obj_id; job.status = status;
    job.nsteps = nsteps;
    rtgr::Counters cnt{0, 0, 0, 0};
    dispatch(sc, p->r_formula, job, cnt);
    if (counters) { counters[0] = cnt.rays; counters[1] = cnt.attempts; counters[2] = cnt.accepted; counters[3] = cnt.rejected; }
    return 0;
}
// The whole frame's tiles in the frame order of rtgr_render_frame, drawn from the shared head `head`;
// pixels are stored straight into the (shared) image `rgb8`.
int shim_render_frame(const rtgr_params* p, const rtgr_object* objs, int n_objs, const rtgr_camera* cam,
                      unsigned long long* head, uint8_t* rgb8, uint64_t* counters) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, objs, n_objs, cam, sc, err)) return -1;
    rtgr::Job job{};
    job.mode = rtgr::JOB_RENDER; job.tile_offset = 0; job.tile_stride = 1; job.queue_scope = 1;
    int64_t count;
    rtgr::tile_selection(cam->ni, cam->nj, 0, 1, job.tiles_x, count);
    job.total = count * (RTGR_TILE_W * RTGR_TILE_H);
    std::vector<int32_t> order;
    if (p->metric == RTGR_KERR_SCHILD) { order = rtgr::tile_order_by_impact(*cam); job.tile_order = order.data(); }
    job.rgb8 = rgb8; job.rgb_stride = 3;
    rtgr::Counters cnt{0, 0, 0, 0};
    SharedQueueSched s{head};
    dispatch_with(sc, p->r_formula, job, cnt, s);
    if (counters) { counters[0] = cnt.rays; counters[1] = cnt.attempts; counters[2] = cnt.accepted; counters[3] = cnt.rejected; }
    return 0;
}
// Host model of rtgr_trace_canvas_frame: the rays of ONE Pixel canvas (shared by the participating processes) are
// drawn from the shared head `head` in the frame order derived from the canvas's contents; rgb is written in place.
int shim_trace_canvas_frame(const rtgr_params* p, const rtgr_object* objs, int n_objs, rtgr_pixel* pixels, int ni, int nj,
                            unsigned long long* head, uint64_t* counters) {
    rtgr::SceneConst sc; std::string err;
    if (!rtgr::build_scene_const(p, objs, n_objs, nullptr, sc, err)) return -1;
    sc.ni = ni; sc.nj = nj;
    rtgr::Job job{};
    job.mode = rtgr::JOB_RENDER; job.tile_offset = 0; job.tile_stride = 1; job.queue_scope = 1;
    int64_t count;
    rtgr::tile_selection(ni, nj, 0, 1, job.tiles_x, count);
    job.total = count * (RTGR_TILE_W * RTGR_TILE_H);
    std::vector<int32_t> order;
    if (p->metric == RTGR_KERR_SCHILD) { order = rtgr::tiles_sorted_by_key(rtgr::tile_impact_keys_pixels(pixels, ni, nj)); job.tile_order = order.data(); }
    double* dpx = reinterpret_cast<double*>(pixels);
    job.pixels_in = dpx; job.rgb_f64 = dpx + 8; job.rgb_stride = 11;
    rtgr::Counters cnt{0, 0, 0, 0};
    SharedQueueSched s{head};
    dispatch_with(sc, p->r_formula, job, cnt, s);
    if (counters) { counters[0] = cnt.rays; counters[1] = cnt.attempts; counters[2] = cnt.accepted; counters[3] = cnt.rejected; }
    return 0;
}
}
