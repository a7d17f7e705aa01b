// oracle/oracle_api.cpp -- TEST INFRASTRUCTURE. C API of liboracle.so: batch runner of the
// restated oracle (planner_oracle.cpp) over independent scenes on T host threads, with the same
// [cycle][scene] array layout the CUDA path and oracle/_ref/libref.so use, plus thin entry points
// to the operator specification (cshare_spec.cpp) for known-answer tests.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>
#include "planner_oracle.h"
#include "world_spec.h"
#include "v2x_oracle.h"

namespace { oracle::MapView g_map; bool g_have_map = false; }

extern "C" {

int oracle_set_map(const dp_map_desc* m) { g_map.d = *m; g_have_map = true; return 0; }   // caller keeps arrays alive

void oracle_default_params(dp_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->vehicle_width = 1.8; p->epsilon = 1e-6; p->pi = 3.14159265358979323846;
    p->road_faraim_max = 60; p->road_faraim_min = 15; p->pre_inter_faraim = 20; p->inter_faraim = 15;
    p->road_remain_distance = 15; p->inter_remain_distance = 5;
    p->lat0 = 23.0; p->lng0 = 113.0; p->k_lat = 1.0 / 110574.0; p->k_lng = 1.0 / 102470.0;
    p->id_more = 8;
}

// hdr[cycles][n], obs_[xy][cycles][n][max_obs], rec[cycles][n], trace[cycles][n] (nullable),
// path_xy[cycles][n][2][200] (nullable), path_ll[cycles][n][2][100] (nullable),
// calls[cycles][n][calls_cap] + n_calls[cycles][n] (nullable), carry_out[n] / last_path_out[n][2][200] (nullable).
// Returns total ub_hits (>= 0) or a negative error.  *seconds = wall time of the compute loop.
long long oracle_run_batch(const dp_params* p, int n, int cycles, int max_obs, const dp_scene_hdr* hdr,
                           const double* ox, const double* oy, dp_plan_record* rec, dp_trace_record* trace,
                           double* path_xy, double* path_ll, ref_call* calls, int32_t* n_calls, int calls_cap,
                           dp_carry* carry_out, double* last_path_out, int exhaustive, int threads,
                           double* seconds, long long* traj_scored) {
    if (!g_have_map) return -1;
    if (threads < 1) threads = 1;
    std::atomic<long long> ub(0), traj(0);
    auto work = [&](int s0, int s1) {
        long long lub = 0, ltraj = 0;
        for (int s = s0; s < s1; ++s) {
            oracle::SceneState st;
            oracle::reset_state(st);
            for (int c = 0; c < cycles; ++c) {
                size_t e = (size_t)c * n + s;
                oracle::CycleOut o{};
                o.rec = rec + e;
                o.trace = trace ? trace + e : nullptr;
                o.path_xy = path_xy ? path_xy + e * 400 : nullptr;
                o.path_ll = path_ll ? path_ll + e * 200 : nullptr;
                o.calls = calls ? calls + e * calls_cap : nullptr;
                o.calls_cap = calls_cap;
                oracle::cycle(g_map, *p, hdr[e], ox + e * max_obs, oy + e * max_obs, st, o, exhaustive != 0);
                if (n_calls) n_calls[e] = o.n_calls;
                lub += o.ub_hits;
                ltraj += o.n_calls;
            }
            if (carry_out) carry_out[s] = st.c;
            if (last_path_out) {
                std::memcpy(last_path_out + (size_t)s * 400, st.last_x, sizeof(st.last_x));
                std::memcpy(last_path_out + (size_t)s * 400 + 200, st.last_y, sizeof(st.last_y));
            }
        }
        ub += lub; traj += ltraj;
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, (int)((long long)n * t / threads), (int)((long long)n * (t + 1) / threads));
        for (auto& x : th) x.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (traj_scored) *traj_scored = traj.load();
    return ub.load();
}

// Same as oracle_run_batch with predicted agent tracks (BASELINE config 5): per cycle and scene the constant-turn-rate parameters
// trk_vx / trk_vy [m per step] and trk_dth [deg per step], [cycles][n][max_obs]; the [T x n_obs] tile of a scene is rolled out from
// its obstacle positions (spec::rollout_ctr) and the junction search runs against it.  No reference counterpart exists.
long long oracle_run_batch_tracks(const dp_params* p, int n, int cycles, int max_obs, const dp_scene_hdr* hdr, const double* ox,
                                  const double* oy, const double* trk_vx, const double* trk_vy, const double* trk_dth, int T,
                                  dp_plan_record* rec, dp_trace_record* trace, double* path_xy, double* path_ll, dp_carry* carry_out,
                                  double* last_path_out, int threads, double* seconds) {
    if (!g_have_map) return -1;
    if (threads < 1) threads = 1;
    std::atomic<long long> ub(0);
    auto work = [&](int s0, int s1) {
        std::vector<double> tx((size_t)T * max_obs), ty((size_t)T * max_obs);
        long long lub = 0;
        for (int s = s0; s < s1; ++s) {
            oracle::SceneState st;
            oracle::reset_state(st);
            for (int c = 0; c < cycles; ++c) {
                size_t e = (size_t)c * n + s;
                const int no = hdr[e].n_obs < max_obs ? hdr[e].n_obs : max_obs;
                for (int o = 0; o < no; ++o)
                    spec::rollout_ctr(ox[e * max_obs + o], oy[e * max_obs + o], trk_vx[e * max_obs + o], trk_vy[e * max_obs + o],
                                      trk_dth[e * max_obs + o], T, tx.data() + o, ty.data() + o, no);
                oracle::CycleOut o{};
                o.rec = rec + e;
                o.trace = trace ? trace + e : nullptr;
                o.path_xy = path_xy ? path_xy + e * 400 : nullptr;
                o.path_ll = path_ll ? path_ll + e * 200 : nullptr;
                o.tile_x = tx.data(); o.tile_y = ty.data(); o.tile_T = T;
                oracle::cycle(g_map, *p, hdr[e], ox + e * max_obs, oy + e * max_obs, st, o, true);
                lub += o.ub_hits;
            }
            if (carry_out) carry_out[s] = st.c;
            if (last_path_out) {
                std::memcpy(last_path_out + (size_t)s * 400, st.last_x, sizeof(st.last_x));
                std::memcpy(last_path_out + (size_t)s * 400 + 200, st.last_y, sizeof(st.last_y));
            }
        }
        ub += lub;
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, (int)((long long)n * t / threads), (int)((long long)n * (t + 1) / threads));
        for (auto& x : th) x.join();
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return ub.load();
}
void oracle_search_obstacle_tile(const double* px, const double* py, int P, const double* tx, const double* ty, int T, int N, double lo,
                                 double hi, dp_search_slot* out) {
    std::vector<spec::P2> path(P);
    for (int i = 0; i < P; ++i) path[i] = spec::P2{px[i], py[i]};
    spec::SearchResult r = spec::search_obstacle_tile(path.data(), P, tx, ty, T, N, lo, hi);
    std::memset(out, 0, sizeof(*out));
    out->dis_lat = r.dis_lat; out->dis_lng = r.dis_lng; out->ob_index = (int16_t)r.ob_index; out->pathid = (uint16_t)r.pathid;
    out->evaluated = 1; out->found = r.found;
}
void oracle_rollout_ctr(double x0, double y0, double vx, double vy, double dth, int T, double* out_x, double* out_y) {
    spec::rollout_ctr(x0, y0, vx, vy, dth, T, out_x, out_y, 1);
}

// hit counters of the right-lane-change sites since the last reset (planner_oracle.h, BR_*)
void oracle_branch_hits(long long* out, int reset) {
    for (int i = 0; i < oracle::BR_COUNT; ++i) {
        out[i] = oracle::g_branch_hits[i].load();
        if (reset) oracle::g_branch_hits[i].store(0);
    }
}

// ---- operator-level entry points (known-answer tests, GPU operator parity) ----
void oracle_search_obstacle(const double* px, const double* py, int P, const double* ox, const double* oy, int N,
                            double lo, double hi, dp_search_slot* out) {
    std::vector<spec::P2> p(P), o(N);
    for (int i = 0; i < P; ++i) p[i] = spec::P2{px[i], py[i]};
    for (int i = 0; i < N; ++i) o[i] = spec::P2{ox[i], oy[i]};
    spec::SearchResult r = spec::search_obstacle(p.data(), P, o.data(), N, lo, hi);
    std::memset(out, 0, sizeof(*out));
    out->dis_lat = r.dis_lat; out->dis_lng = r.dis_lng; out->ob_index = (int16_t)r.ob_index;
    out->pathid = (uint16_t)r.pathid; out->evaluated = 1; out->found = r.found;
}
void oracle_create_new_path(const double* px, const double* py, int P, double d, double* ox, double* oy) {
    std::vector<spec::P2> p(P), o(P);
    for (int i = 0; i < P; ++i) p[i] = spec::P2{px[i], py[i]};
    spec::create_new_path(p.data(), P, d, o.data());
    for (int i = 0; i < P; ++i) { ox[i] = o[i].x; oy[i] = o[i].y; }
}
void oracle_bezier(const double* poses6, double* out_xy /*[2][n]*/, int n) {
    std::vector<spec::P2> o(n);
    spec::bezier_planning(spec::P3{poses6[0], poses6[1], poses6[2]}, spec::P3{poses6[3], poses6[4], poses6[5]}, o.data(), n);
    for (int i = 0; i < n; ++i) { out_xy[i] = o[i].x; out_xy[n + i] = o[i].y; }
}
int oracle_nearest_id(double qx, double qy, const double* px, const double* py, int n) {
    std::vector<spec::P2> p(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) p[i] = spec::P2{px[i], py[i]};
    return spec::nearest_id(spec::P2{qx, qy}, p.data(), n);
}
void oracle_mean_points(const double* px, const double* py, int n_in, double* out_xy /*[2][n_out]*/, int n_out) {
    std::vector<spec::P2> p(n_in > 0 ? n_in : 1), o(n_out);
    for (int i = 0; i < n_in; ++i) p[i] = spec::P2{px[i], py[i]};
    spec::mean_points(p.data(), n_in, o.data(), n_out);
    for (int i = 0; i < n_out; ++i) { out_xy[i] = o[i].x; out_xy[n_out + i] = o[i].y; }
}
// dense candidate sweep (BASELINE config 3): CPU statement of dp_score_candidates
int oracle_score_candidates(const double* bx, const double* by, int n_base, const double* offset, const int32_t* n_pts, int n_cand,
                            const double* ox, const double* oy, const double* dvx, const double* dvy, int n_obs, double lo,
                            double hi, double clear_dis, double* out_dis_lng) {
    std::vector<spec::P2> base(n_base), cand(n_base), o(n_obs), dv(n_obs);
    for (int i = 0; i < n_base; ++i) base[i] = spec::P2{bx[i], by[i]};
    for (int i = 0; i < n_obs; ++i) { o[i] = spec::P2{ox[i], oy[i]}; dv[i] = spec::P2{dvx ? dvx[i] : 0.0, dvy ? dvy[i] : 0.0}; }
    int best = -1;
    for (int c = 0; c < n_cand; ++c) {
        int P = n_pts[c] < n_base ? n_pts[c] : n_base;
        double dis = spec::NOT_FOUND;
        if (P >= 2) {
            spec::create_new_path(base.data(), P, offset[c], cand.data());
            dis = spec::search_obstacle_tracks(cand.data(), P, o.data(), dv.data(), n_obs, lo, hi).dis_lng;
        }
        if (out_dis_lng) out_dis_lng[c] = dis;
        if (best < 0 && dis > clear_dis) best = c;      // first feasible == lowest feasible index
    }
    return best;
}
void oracle_sincos_deg(double a, double* c, double* s) { spec::spec_sincos_deg(a, c, s); }
double oracle_atan(double z) { return spec::spec_atan(z); }
double oracle_calc_global_dir(double ax, double ay, double bx, double by) {
    return spec::calc_global_dir({ax, ay}, {bx, by}, 1e-6, 3.14159265358979323846);
}
double oracle_lat_dis(double qx, double qy, double ax, double ay, double bx, double by) {
    return spec::lat_dis({qx, qy}, {ax, ay}, {bx, by}, 1e-6);
}
double oracle_calc_distance(double ax, double ay, double bx, double by) { return spec::calc_distance({ax, ay}, {bx, by}); }

// ---- closed-loop episodes and output frames (world_spec.h) ----
void oracle_world_default_params(dp_world_params* p) { oracle::world_default_params(p); }

// one world step for n scenes from host arrays; rec nullable (place + localise only); last_path[n][2][200] (ignored when rec is null)
int oracle_world_step(const dp_params* p, const dp_world_params* wp, int n, int max_obs, dp_scene_hdr* hdr, dp_agent* agents, double* ox,
                      double* oy, const dp_plan_record* rec, const double* last_path) {
    if (!g_have_map) return -1;
    for (int s = 0; s < n; ++s)
        oracle::world_step(g_map, *p, *wp, hdr[s], agents + (size_t)s * max_obs, ox + (size_t)s * max_obs, oy + (size_t)s * max_obs,
                           rec ? rec + s : nullptr, last_path ? last_path + (size_t)s * 400 : nullptr,
                           last_path ? last_path + (size_t)s * 400 + 200 : nullptr);
    return 0;
}

// `cycles` x (cycle, world step) per scene from the world in hdr[n] / agents[n][max_obs] (updated in place, like ox / oy[n][max_obs]).
// rec[cycles][n]; hdr_log[cycles][n], obs_log_[xy][cycles][n][max_obs], path_xy[cycles][n][2][200] nullable.
long long oracle_run_closed_loop(const dp_params* p, const dp_world_params* wp, int n, int cycles, int max_obs, dp_scene_hdr* hdr,
                                 dp_agent* agents, double* ox, double* oy, dp_plan_record* rec, dp_scene_hdr* hdr_log, double* obs_log_x,
                                 double* obs_log_y, double* path_xy, dp_carry* carry_out, double* last_path_out, int threads,
                                 int32_t* ub_scene /* nullable [n]: reference-UB sites hit by each scene */) {
    if (!g_have_map) return -1;
    if (threads < 1) threads = 1;
    std::atomic<long long> ub(0);
    auto work = [&](int s0, int s1) {
        long long lub = 0;
        for (int s = s0; s < s1; ++s) {
            const long long lub0 = lub;
            oracle::SceneState st;
            oracle::reset_state(st);
            dp_agent* ag = agents + (size_t)s * max_obs;
            double* sx = ox + (size_t)s * max_obs; double* sy = oy + (size_t)s * max_obs;
            oracle::world_step(g_map, *p, *wp, hdr[s], ag, sx, sy, nullptr, nullptr, nullptr);
            for (int c = 0; c < cycles; ++c) {
                const size_t e = (size_t)c * n + s;
                if (hdr_log) hdr_log[e] = hdr[s];
                if (obs_log_x) std::memcpy(obs_log_x + e * max_obs, sx, sizeof(double) * max_obs);
                if (obs_log_y) std::memcpy(obs_log_y + e * max_obs, sy, sizeof(double) * max_obs);
                oracle::CycleOut o{};
                o.rec = rec + e;
                o.path_xy = path_xy ? path_xy + e * 400 : nullptr;
                oracle::cycle(g_map, *p, hdr[s], sx, sy, st, o, false);
                lub += o.ub_hits;
                oracle::world_step(g_map, *p, *wp, hdr[s], ag, sx, sy, rec + e, st.last_x, st.last_y);
            }
            if (ub_scene) ub_scene[s] = (int32_t)(lub - lub0);
            if (carry_out) carry_out[s] = st.c;
            if (last_path_out) {
                std::memcpy(last_path_out + (size_t)s * 400, st.last_x, sizeof(st.last_x));
                std::memcpy(last_path_out + (size_t)s * 400 + 200, st.last_y, sizeof(st.last_y));
            }
        }
        ub += lub;
    };
    if (threads == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(work, (int)((long long)n * t / threads), (int)((long long)n * (t + 1) / threads));
        for (auto& t : th) t.join();
    }
    return ub.load();
}

// frames of n scenes: rec[n], path_xy[n][2][200] = road_points of the cycle; ctrl / status nullable
void oracle_pack_frames(const dp_params* p, int n, const dp_plan_record* rec, const double* path_xy, dp_ctrl_frame* ctrl,
                        dp_status_frame* status) {
    for (int s = 0; s < n; ++s)
        oracle::pack_frames(*p, rec[s], path_xy + (size_t)s * 400, path_xy + (size_t)s * 400 + 200, ctrl ? ctrl + s : nullptr,
                            status ? status + s : nullptr);
}
// V2X event handlers of n scenes (v2x_oracle.h)
int oracle_v2x_event(const dp_params* p, int n, const dp_scene_hdr* hdr, const dp_v2x_data* v2x, const double* wp_lat, const double* wp_lng,
                     int mode, dp_v2x_flags* out) {
    if (!g_have_map) return -1;
    for (int s = 0; s < n; ++s) oracle::v2x_event(g_map, *p, hdr[s], v2x[s], wp_lat, wp_lng, mode, out + s);
    return 0;
}
void oracle_v2x_apply(int n, const dp_v2x_flags* flags, dp_plan_record* rec) {
    for (int s = 0; s < n; ++s) oracle::v2x_apply(flags[s], rec[s]);
}
int oracle_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(dp_scene_hdr);
        case 1: return (int)sizeof(dp_plan_record);
        case 2: return (int)sizeof(dp_carry);
        case 3: return (int)sizeof(ref_call);
        case 4: return (int)sizeof(dp_trace_record);
        case 5: return (int)sizeof(dp_params);
        case 6: return (int)sizeof(dp_map_desc);
        case 7: return (int)sizeof(dp_ctrl_frame);
        case 8: return (int)sizeof(dp_status_frame);
        case 9: return (int)sizeof(dp_agent);
        case 10: return (int)sizeof(dp_world_params);
        case 11: return (int)sizeof(dp_v2x_data);
        case 12: return (int)sizeof(dp_v2x_flags);
    }
    return -1;
}
}  // extern "C"
