// oracle/planner_oracle.cpp -- TEST INFRASTRUCTURE (CPU oracle); contract in planner_oracle.h.
// Every block cites the reference lines it restates.  Compile with -ffp-contract=off.
#include "planner_oracle.h"
#include <cmath>
#include <cstring>

#include <atomic>

namespace oracle {
using spec::P2;
using spec::P3;
typedef std::vector<P2> Path;

// per-branch hit counters of the right-lane-change sites (tests assert that the directed scene families reach them)
std::atomic<long long> g_branch_hits[BR_COUNT];
static inline void hit(int b) { g_branch_hits[b].fetch_add(1, std::memory_order_relaxed); }

void reset_state(SceneState& s) {
    std::memset(&s, 0, sizeof(s));
    s.c.behavior = 1;            // Decision.cpp:15
    s.c.velocity_expect = 10;    // Decision.cpp:16
    s.c.his_behavior = 1;        // Decision.cpp:25
    s.c.plan_his_behavior = 1;   // Planning.cpp:10
}

namespace {

struct Gap { double dis_lat, dis_lng; bool flag; int pathid; int ob; };   // Path_Obs, zero = memset state
struct Beh { int behavior, target, light; bool lanechg, obsavoid; int dlg; };   // Behavior_Dec

struct Ctx {
    const MapView& m;
    const dp_params& p;
    const dp_scene_hdr& h;
    const P2* obs;
    int n_obs;
    CycleOut& out;
    bool exhaustive;
    int n_traj = 0;
    long pts = 0;
};

// one CShare::SearchObstacle evaluation + call log entry
spec::SearchResult search(Ctx& k, const Path& path, double lo, double hi, dp_search_slot* slot) {
    spec::SearchResult r = spec::search_obstacle(path.data(), (int)path.size(), k.obs, k.n_obs, lo, hi);
    if (k.out.calls && k.out.n_calls < k.out.calls_cap) {
        ref_call& c = k.out.calls[k.out.n_calls];
        std::memset(&c, 0, sizeof(c));
        c.lat_min = lo; c.lat_max = hi; c.dis_lat = r.dis_lat; c.dis_lng = r.dis_lng;
        c.n_path = (int32_t)path.size(); c.ob_index = (int16_t)r.ob_index; c.pathid = (uint16_t)r.pathid;
        c.found = r.found;
    }
    ++k.out.n_calls;
    ++k.n_traj;
    k.pts += (long)path.size();
    if (slot) {
        slot->dis_lat = r.dis_lat; slot->dis_lng = r.dis_lng; slot->ob_index = (int16_t)r.ob_index;
        slot->pathid = (uint16_t)r.pathid; slot->evaluated = 1; slot->found = r.found;
    }
    return r;
}

Path offset_path(const Path& p, double d) {
    Path o(p.size());
    spec::create_new_path(p.data(), (int)p.size(), d, o.data());
    return o;
}

// map point read with the reference's out-of-range index (== size) defined as "last point"
P2 map_pt(Ctx& k, int gl, int i) {
    int n = k.m.lane_size(gl);
    if (i >= n) { ++k.out.ub_hits; i = n - 1; }
    if (i < 0) { ++k.out.ub_hits; i = 0; }
    return k.m.pt(gl, i);
}

// Decision.cpp:581-596 / 611-622 / 649-660: forward 120-point and backward 40-point slices
void load_lane_paths(Ctx& k, int gl, int id, Path& fwd, Path& rear) {
    int n = k.m.lane_size(gl);
    int more = k.p.id_more;
    int a = std::min(n, id + more), b = std::min(n, id + 120 + more);
    for (int i = a; i < b; ++i) fwd.push_back(map_pt(k, gl, i));
    int lo = std::max(0, id + more - 40);
    // the reference starts this loop at index == size when id + ID_MORE >= size (out-of-range read,
    // Decision.cpp:590-594); defined here as: start from the last valid point, same point count
    int start = a;
    if (start >= n) { start = n - 1; if (a > lo) ++k.out.ub_hits; }
    for (int t = 0; t < a - lo; ++t) rear.push_back(map_pt(k, gl, start - t));
}

// Decision.cpp:1179-1187 and siblings: arclength from Id_CurLane while cond(attr[i+1])
template <class Cond>
double lanechg_run(Ctx& k, int gl, int id, Cond cond) {
    int n = k.m.lane_size(gl);
    double dis = 0;
    for (int i = id; (i < n - 1) && cond(k.m.attr(gl, i + 1)); ++i) dis += spec::calc_distance(k.m.pt(gl, i), k.m.pt(gl, i + 1));
    return dis;
}

// Decision.cpp:498-538
int navi_times(const uint16_t* out, int lane, int dir) {
    int times = 5;
    for (int i = 0; i < DP_LANESUM && out[i] != 0; ++i) {
        int t = (dir == 1) ? lane - out[i] : out[i] - lane;
        if (t < times) times = t;
    }
    return (unsigned char)times;   // BYTE return
}

// ------------------------------------------------------------------------------------------------
// SegmentDecision (Decision.cpp:216-315)
// ------------------------------------------------------------------------------------------------
void segment_decision(Ctx& k, SceneState& st, Path& refpath) {
    const dp_scene_hdr& h = k.h;
    dp_carry& c = st.c;
    dp_trace_record* tr = k.out.trace;
    const int road = h.road_num, lane = h.lane_num;
    const int gl = k.m.lane_index(road, lane);
    const int id = (uint16_t)h.id[lane - 1];            // WORD Id_CurLane (Decision.cpp:561)
    const int id_sum = k.m.lane_size(gl);
    const int lane_sum = k.m.lanes_of(road);
    const double Vw = k.p.vehicle_width;

    // ---- Nav_LaneChange (Decision.cpp:685-738) ----
    unsigned navi = 4, navi_t = 0;
    for (int i = 0; i < DP_LANESUM && h.out_lane_no[i] != 0; ++i)
        if (lane == h.out_lane_no[i]) { navi = 0; break; }
    int out_min = h.out_lane_no[0], out_max = 1;
    for (int i = 0; i < DP_LANESUM; ++i) if (h.out_lane_no[i] > out_max) out_max = h.out_lane_no[i];
    if (navi == 4) {
        if (lane < out_min) { navi = 2; navi_t = navi_times(h.out_lane_no, lane, 2); }
        else if (lane > out_max) { navi = 1; navi_t = navi_times(h.out_lane_no, lane, 1); }
        else navi = 0;
    }

    // ---- LoadRefPath (Decision.cpp:553-673) ----
    Path F, R, LF, LR, RF, RR;
    const int lanechg = k.m.attr(gl, std::min(id, id_sum - 1));
    const double W = k.m.width(gl, std::min(id, id_sum - 1)) / 100.0;
    load_lane_paths(k, gl, id, F, R);
    if (lanechg == 1 || lanechg == 3) {
        if (lane > 1) {
            int idl = (uint16_t)h.id[lane - 2], gll = gl - 1, nl = k.m.lane_size(gll);
            if (idl > 0 && idl < nl) load_lane_paths(k, gll, idl, LF, LR);
        } else {
            LF = offset_path(F, -1 * W);
            LR = offset_path(R, -1 * W);
        }
    }
    if (lanechg == 2) {                                  // never for 3
        if (lane < lane_sum) {
            int idr = (uint16_t)h.id[lane], glr = gl + 1, nr = k.m.lane_size(glr);
            if (idr > 0 && idr < nr) load_lane_paths(k, glr, idr, RF, RR);
        } else {
            RF = offset_path(F, W);
            RR = offset_path(R, W);
        }
    }

    // ---- AroundObstacle (Decision.cpp:759-881) ----
    Gap g[6];
    std::memset(g, 0, sizeof(g));
    const Path* paths[6] = {&F, &R, &LF, &LR, &RF, &RR};
    const double lo[6] = {-0.5 * Vw, -0.5 * Vw, -0.5 * Vw, -0.5 * Vw, -0.5 * W, -0.5 * W};
    const double hi[6] = {0.5 * Vw, 0.5 * Vw, 0.5 * W, 0.5 * W, 0.5 * Vw, 0.5 * Vw};
    for (int r = 0; r < 6; ++r) {
        if (paths[r]->size() != 0) {
            spec::SearchResult s = search(k, *paths[r], lo[r], hi[r], tr ? &tr->region[r] : nullptr);
            g[r] = Gap{s.dis_lat, s.dis_lng, s.found, s.pathid, s.ob_index};
        }
    }
    const double gF = g[0].dis_lng, gLF = g[2].dis_lng, gLR = g[3].dis_lng, gRF = g[4].dis_lng, gRR = g[5].dis_lng;
    if (tr) { tr->width_curlane = W; tr->navi_lanechg = navi; tr->navi_lanechg_times = navi_t; }

    // ---- BehaviorDecision (Decision.cpp:898-1773) ----
    Beh cur{c.behavior, c.target_lanenum, c.light_status, c.lanechg_status != 0, c.obsavoid_status != 0, c.behavior_to_dlg};
    const Beh his{c.his_behavior, c.his_target_lanenum, c.his_light_status, false, false, 0};
    int lane_cur = lane;                                 // LaneNum_Cur is assigned at Decision.cpp:1712
    int z_light = c.light_status;                        // member written at Decision.cpp:1193-1206 (dead, quirk 13)
    const double period = h.period_ms;
    int sweep_pick = -1;
    auto keep = [&](bool reset_status) {
        cur.behavior = 1; cur.target = lane_cur;
        if (reset_status) cur.lanechg = false;
    };
    const int lanechg_map = lanechg;
    const int K = [&] { int n = 0; while (n < DP_MAX_SWEEP && (double)n < (W - Vw) / 0.6) ++n; return n; }();

    if (lanechg_map == 0) {                              // :920-1010
        if (gF < 15) {
            c.no_obsavoid_time = 0;
            c.obsavoid_time++;
            if (c.obsavoid_time > 2) {
                // candidates the reference never reaches after its `break` are scored "quietly"
                // (not logged, not counted) only when an exhaustive trace was asked for
                const bool fill = k.exhaustive && tr;
                auto quiet = [&](const Path& cand, dp_search_slot* slot) {
                    spec::SearchResult s = spec::search_obstacle(cand.data(), (int)cand.size(), k.obs, k.n_obs, -0.5 * Vw, 0.5 * Vw);
                    slot->dis_lat = s.dis_lat; slot->dis_lng = s.dis_lng; slot->ob_index = (int16_t)s.ob_index;
                    slot->pathid = (uint16_t)s.pathid; slot->evaluated = 2; slot->found = s.found;
                };
                bool left = false;
                for (int i = 0; i < K; ++i) {            // :940-954
                    Path cand = offset_path(F, -0.3 * i);
                    if (left) { if (!fill) break; quiet(cand, &tr->sweep[i]); continue; }
                    spec::SearchResult s = search(k, cand, -0.5 * Vw, 0.5 * Vw, tr ? &tr->sweep[i] : nullptr);
                    if (s.dis_lng > 25) {
                        cur.behavior = 4; cur.target = lane_cur; cur.light = 1; cur.obsavoid = true; cur.dlg = 11;
                        left = true; sweep_pick = i;
                    }
                }
                bool right = false;
                for (int i = 0; i < K; ++i) {            // :959-973
                    Path cand = offset_path(F, 0.3 * i);
                    if (left || right) { if (!fill) break; quiet(cand, &tr->sweep[DP_MAX_SWEEP + i]); continue; }
                    spec::SearchResult s = search(k, cand, -0.5 * Vw, 0.5 * Vw, tr ? &tr->sweep[DP_MAX_SWEEP + i] : nullptr);
                    if (s.dis_lng > 25) {
                        cur.behavior = 5; cur.target = lane_cur; cur.light = 2; cur.obsavoid = true; cur.dlg = 12;
                        right = true; sweep_pick = K + i;
                    }
                }
            } else {
                cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1;
            }
        } else {
            if (c.obsavoid_status == 0) {                // z_segment_obsavoid_status :987
                cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1;
            } else {
                c.no_obsavoid_time++;
                if (c.no_obsavoid_time > 3) {
                    cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1; cur.obsavoid = false;
                }
            }
            cur.dlg = 1;
        }
    } else {                                             // :1012-1772
        c.no_obsavoid_time = 0;
        c.obsavoid_time = 0;
        if (cur.lanechg == 0) {
            if (navi != 0) {                             // :1021-1144
                if (navi == 1) {
                    if (lanechg_map == 1 || lanechg_map == 3) {
                        cur.dlg = 2;
                        if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                        c.leftlight_time += period;
                        if (((gLF > gF + 10) || (gLF > 40)) && gLR > 15 && c.leftlight_time > 2000) {
                            cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                        } else keep(true);
                    } else { keep(true); cur.dlg = 4; }
                } else if (navi == 2) {
                    if (lanechg_map == 2 || lanechg_map == 3) {
                        cur.dlg = 3;
                        if (cur.light != 2) { cur.lanechg = true /* = 2, quirk 3 */; c.rightlight_time = 0; }
                        c.rightlight_time += period;
                        if (((gRF > gF + 10) || (gRF > 40)) && gRR > 15 && c.rightlight_time >= 2000) {
                            hit(BR_B3_NAV_1108);
                            cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                        } else keep(true);
                    } else { keep(true); cur.dlg = 4; }
                }
            } else {                                     // :1146-1757
                if (gF < (2 * 10 + 5)) {
                    c.frontobs_time++;
                    if (c.frontobs_time > 2) {
                        c.frontobs_time = 3;
                        auto out_has = [&](int l) { bool f = false; for (int i = 0; i < DP_LANESUM && h.out_lane_no[i] != 0; ++i) if (h.out_lane_no[i] == l) f = true; return f; };
                        auto bit0 = [](int a) { return (a & 1) != 0; };        // "& 0x01 == 0x01" and "& 0x02 == 0x02" (quirk 1)
                        if (lanechg_map == 1) {          // :1157-1294
                            if (lane_cur > 1) {
                                cur.dlg = 5;
                                bool no_back = !out_has(lane_cur - 1);
                                bool chg = false;
                                if (no_back) {
                                    double d1 = lanechg_run(k, gl, id, [](int a) { return a == 1; });
                                    if (d1 > 60) {
                                        chg = true;
                                        if (z_light != 1) { z_light = 1; c.leftlight_time = 0; }
                                        c.leftlight_time += period;
                                        if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                    } else z_light = 0;
                                } else {
                                    double d1 = lanechg_run(k, gl, id, bit0);
                                    if (d1 > 15) {
                                        chg = true;
                                        if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                                        c.leftlight_time += period;
                                        if (c.leftlight_time > 2000) c.leftlight_time = 2100;
                                    } else cur.light = 0;
                                }
                                if (chg && gLF > gF + 10 && gLR > 10 && c.leftlight_time > 1500) {
                                    c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                } else keep(true);
                            } else keep(true);
                        } else if (lanechg_map == 2) {   // :1296-1424
                            if (lane_cur < lane_sum) {
                                cur.dlg = 6;
                                bool no_back = !out_has(lane_cur - 1);          // sic: lane-1 (:1307)
                                bool chg = false;
                                double d1 = lanechg_run(k, gl, id, bit0);
                                if (no_back) {
                                    if (d1 > 50) {
                                        chg = true;
                                        if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                        c.leftlight_time += period;
                                        if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                    }
                                } else {
                                    if (d1 > 10) {
                                        chg = true;
                                        if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                                        c.leftlight_time += period;
                                        if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                    }
                                }
                                if (chg && gRF > gF + 10 && gRR > 10 && c.leftlight_time > 1500) {
                                    hit(BR_B3_OBS_1382);
                                    c.frontobs_time = 0; cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                                } else keep(true);
                            } else keep(true);
                        } else if (lanechg_map == 3) {   // :1426-1738
                            bool nb_left = !out_has(lane_cur - 1), nb_right = !out_has(lane_cur + 1);
                            bool left_ok = false, right_ok = false;
                            if (lane_cur > 1) {
                                if (nb_left) { if (lanechg_run(k, gl, id, bit0) > 50) left_ok = true; }
                                else { if (lanechg_run(k, gl, id, [](int) { return false; }) > 10) left_ok = true; }   // "& 0x01 != 0x01" == 0
                            }
                            if (lane_cur < lane_sum) {
                                if (nb_right) { if (lanechg_run(k, gl, id, bit0) > 50) right_ok = true; }
                                else { if (lanechg_run(k, gl, id, bit0) > 10) right_ok = true; }
                            }
                            auto tick = [&] { c.leftlight_time += period; if (c.leftlight_time > 2000) c.leftlight_time = 2000; };
                            if (left_ok && !nb_left) {                               // :1545-1594
                                if (cur.lanechg != 1) { cur.light = 1; c.leftlight_time = 0; }
                                tick();
                                if (gLF > gF + 10 && gLR > 10 && c.leftlight_time > 2000) {
                                    c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                } else keep(true);
                            } else if (right_ok && !nb_right) {                      // :1596-1636
                                hit(BR_ENTER_1596);
                                if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                tick();
                                if (gRF > gF + 10) {
                                    if (gRR > 10 && c.leftlight_time > 2000) {
                                        hit(BR_B3_BOTH_1618);
                                        c.frontobs_time = 0; cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                                    } else keep(false);
                                }
                            } else if (left_ok) {                                    // :1638-1686
                                if (cur.light != 1) { cur.lanechg = true /* quirk 3 */; c.leftlight_time = 0; }
                                tick();
                                if (gLF > gF + 10 && gLR > 10 && c.leftlight_time > 2000) {
                                    c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                } else keep(false);
                            } else if (right_ok) {                                   // :1688-1730
                                hit(BR_ENTER_1688);
                                if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                tick();
                                if (gRF > gF + 10) {
                                    if (gRR > 10 && c.leftlight_time > 2000) {
                                        hit(BR_B3_BOTH_1711);
                                        c.frontobs_time = 0; cur.behavior = 3;
                                        lane_cur = 1; cur.target = 1;                // "= LaneNum_Cur = 1" (quirk 2)
                                        cur.lanechg = true;
                                    } else keep(false);
                                }
                            } else keep(true);
                        }
                    } else keep(true);                   // :1741-1746
                } else {                                 // :1749-1756
                    c.frontobs_time = 0; cur.dlg = 8; keep(true);
                }
            }
        } else {                                         // lanechg_status == 1, :1760-1771
            cur.dlg = 9;
            if (cur.target == lane_cur) { cur.lanechg = false; cur.light = 0; }
            cur.behavior = his.behavior; cur.target = his.target; cur.light = his.light;
        }
    }
    (void)z_light;

    // ---- SpeedDecision / RefPath / write-back (Decision.cpp:1781-1816, 307-313) ----
    c.velocity_expect = (cur.behavior == 4 || cur.behavior == 5) ? 5 : 10;
    refpath = (cur.behavior == 2) ? LF : (cur.behavior == 3) ? RF : F;
    c.behavior = (uint16_t)cur.behavior; c.light_status = (uint16_t)cur.light; c.target_lanenum = (uint16_t)cur.target;
    c.lanechg_status = cur.lanechg; c.obsavoid_status = cur.obsavoid; c.behavior_to_dlg = (uint16_t)cur.dlg;
    c.target_roadnum = h.road_num;
    k.out.rec->sweep_index = (int16_t)sweep_pick;
}

// ------------------------------------------------------------------------------------------------
// PreStubDecision / StubDecision (Decision.cpp:323-402, 409-486)
// ------------------------------------------------------------------------------------------------
void junction_decision(Ctx& k, SceneState& st, Path& refpath) {
    const dp_scene_hdr& h = k.h;
    dp_carry& c = st.c;
    Path F;
    const int gc = (h.conn >= 0 && h.conn < k.m.d.n_conn) ? k.m.d.conn[h.conn].lane : -1;
    const int n_inter = gc >= 0 ? k.m.lane_size(gc) : 0;
    if (h.pos == 1) {
        int gl = k.m.lane_index(h.road_num, h.lane_num);
        int id = h.id[h.lane_num - 1], n = k.m.lane_size(gl);
        for (int i = (uint16_t)id; i < n; ++i) F.push_back(k.m.pt(gl, i));         // :352-358
        for (int j = 0; j < n_inter; ++j) F.push_back(k.m.pt(gc, j));               // :361-367
    } else {
        int id = h.id[h.last_lanenum - 1];
        for (int j = (uint16_t)id; j < n_inter; ++j) F.push_back(k.m.pt(gc, j));    // :438-444
        int gl = k.m.lane_index(h.road_num, h.lane_num), n = k.m.lane_size(gl);
        for (int i = 0; i < std::min(60, n); ++i) F.push_back(k.m.pt(gl, i));       // :446-452
    }
    const double Vw = k.p.vehicle_width;
    spec::SearchResult s;
    if (k.out.tile_x && k.out.tile_T > 0) {                 // predicted tracks: the same call against the scene's track tile
        s = spec::search_obstacle_tile(F.data(), (int)F.size(), k.out.tile_x, k.out.tile_y, k.out.tile_T, k.n_obs, -0.5 * Vw, 0.5 * Vw);
        ++k.out.n_calls; ++k.n_traj; k.pts += (long)F.size();
        if (k.out.trace) {
            dp_search_slot* slot = &k.out.trace->junction;
            slot->dis_lat = s.dis_lat; slot->dis_lng = s.dis_lng; slot->ob_index = (int16_t)s.ob_index;
            slot->pathid = (uint16_t)s.pathid; slot->evaluated = 1; slot->found = s.found;
        }
    } else s = search(k, F, -0.5 * Vw, 0.5 * Vw, k.out.trace ? &k.out.trace->junction : nullptr);   // :370 / :455
    if (s.dis_lng < 13) {
        double v = s.dis_lng - 3;
        c.velocity_expect = v > 0 ? v : 0;              // max(dis_lng - 3, 0)
        c.behavior_to_dlg = 13;
    } else {
        c.velocity_expect = 10;
        c.behavior_to_dlg = 1;
    }
    c.light_status = (h.stub_attribute == 3) ? 1 : h.stub_attribute;                // :385-392
    c.behavior = 1;
    c.target_roadnum = h.road_num;
    c.target_lanenum = h.lane_num;
    refpath = F;
    k.out.rec->sweep_index = -1;
}

// Planning.cpp:686-709
double get_lat_dis(P2 cur, P2 pt, P2 nx, double eps) { return spec::lat_dis(cur, pt, nx, eps); }

// Planning.cpp:719-750 (libm atan, as the reference)
double get_road_angle(P2 a, P2 b, double eps, double pi) {
    double angle;
    if (std::fabs(b.x - a.x) < eps && std::fabs(b.y - a.y) < eps) angle = 0;
    else if (std::fabs(b.x - a.x) < eps) angle = (b.y > a.y) ? pi / 2 : 3 * pi / 2;
    else {
        angle = std::atan((b.y - a.y) / (b.x - a.x));
        if (b.x < a.x) angle = angle + pi;
        else if ((b.x > a.x) && (b.y < a.y)) angle = angle + 2 * pi;
    }
    return angle * 180 / pi;
}

// Planning.cpp:760-786
double get_angle_err(double d1, double d2) {
    double e = d2 - d1;
    if (d1 < 180) e = (d2 - d1 <= 180) ? d2 - d1 : d2 - d1 - 360;
    else if (d1 >= 180) e = (d2 - d1 > -180) ? d2 - d1 : d2 - d1 + 360;
    return e;
}

// ------------------------------------------------------------------------------------------------
// one iteration of CPlanningThread (Planning.cpp:64-226)
// ------------------------------------------------------------------------------------------------
void planning_cycle(Ctx& k, SceneState& st, const Path& refpath) {
    const dp_scene_hdr& h = k.h;
    const dp_params& p = k.p;
    dp_carry& c = st.c;
    dp_plan_record& r = *k.out.rec;
    const int pos = h.pos;
    const int d_behavior = c.behavior, d_target = c.target_lanenum;
    const double v_exp = c.velocity_expect;

    // ---- Calculate_aim_dis (Planning.cpp:242-290): FLOAT faraim_dis ----
    float faraim = 0;
    if (pos == 0) {
        faraim = (float)((h.velocity / 3.6) * 5 + 4);
        if (faraim > p.road_faraim_max) faraim = (float)p.road_faraim_max;
        else if (faraim < p.road_faraim_min) faraim = (float)p.road_faraim_min;
    } else if (pos == 1) faraim = (float)p.pre_inter_faraim;
    else if (pos == 2) faraim = (float)p.inter_faraim;
    if (k.out.trace) k.out.trace->faraim_dis = faraim;

    // ---- SearchAimPoint (Planning.cpp:303-583) ----
    auto walk_lane = [&](int gl_walk, int from, int to_excl, int gl_fb, int fb_idx, int fb_id) {
        double sum = 0;
        int n_walk = (gl_walk >= 0 && gl_walk < k.m.d.n_lanes) ? k.m.lane_size(gl_walk) : 0;
        for (int i = from; i < to_excl; ++i) {
            if (i + 1 >= n_walk || i < 0) { ++k.out.ub_hits; break; }
            P2 a = k.m.pt(gl_walk, i), b = k.m.pt(gl_walk, i + 1);
            sum += std::sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
            if (sum - 4 > faraim) {
                c.aim_x = a.x; c.aim_y = a.y; c.aim_dir = k.m.dir(gl_walk, i); c.aim_id = i;
                break;
            } else {
                int nf = k.m.lane_size(gl_fb);
                int fi = fb_idx;
                if (fi < 0 || fi >= nf) { ++k.out.ub_hits; fi = fi < 0 ? 0 : nf - 1; }
                P2 f = k.m.pt(gl_fb, fi);
                c.aim_x = f.x; c.aim_y = f.y; c.aim_dir = k.m.dir(gl_fb, fi); c.aim_id = fb_id;
            }
        }
    };
    if (pos == 0) {
        const int road = h.road_num, lane = h.lane_num;
        const int gl = k.m.lane_index(road, lane);
        const int lane_sum = k.m.lanes_of(road);
        const int cur_id = h.id[lane - 1], cur_sum = k.m.lane_size(gl);
        int left_id = 0, left_sum = 0, right_id = 0, right_sum = 0;
        if (lane > 1) { left_id = h.id[lane - 2]; left_sum = k.m.lane_size(gl - 1); }
        if (lane < lane_sum) { right_id = h.id[lane]; right_sum = k.m.lane_size(gl + 1); }
        if (d_target == lane) {
            if (d_behavior == 1) walk_lane(gl, cur_id, cur_sum - 1, gl, cur_sum - 1, cur_sum - 1);   // :404-435
        } else {
            if (d_behavior == 2) {                      // :443-471 (quirk 7)
                if (lane > 1) walk_lane(gl - 1, left_id, left_sum - 1, gl, left_sum - 2, left_sum - 1);
            } else if (d_behavior == 3) {               // :473-501 (quirk 7: bound is leftpoint_sum)
                hit(BR_AIM_RIGHT_473);
                if (lane < lane_sum && right_id < left_sum - 1) hit(BR_AIM_RIGHT_WALK);
                if (lane < lane_sum) walk_lane(gl + 1, right_id, left_sum - 1, gl + 1, right_sum - 1, right_sum - 1);
                else if (right_id < left_sum - 1) ++k.out.ub_hits;   // reference indexes a lane that does not exist
            }
        }
    } else if (pos == 1 || pos == 2) {                  // :505-578
        double sum = 0;
        const int n = (int)refpath.size();
        for (int i = 0; i < n - 1; ++i) {
            sum += spec::calc_distance(refpath[i], refpath[i + 1]);
            if ((sum - 4) > faraim) {
                c.aim_x = refpath[i].x; c.aim_y = refpath[i].y;
                // "i < size() - 4" is an unsigned compare in the reference (size_t)
                if ((size_t)i < refpath.size() - 4) c.aim_dir = spec::calc_global_dir(refpath[i], refpath[i + 2], p.epsilon, p.pi);
                else {
                    int a = i - 2; if (a < 0) { ++k.out.ub_hits; a = 0; }
                    c.aim_dir = spec::calc_global_dir(refpath[a], refpath[i], p.epsilon, p.pi);
                }
                c.aim_id = i;
                break;
            } else {
                int a = n - 3; if (a < 0) { ++k.out.ub_hits; a = 0; }
                c.aim_x = refpath[n - 1].x; c.aim_y = refpath[n - 1].y;
                c.aim_dir = spec::calc_global_dir(refpath[a], refpath[n - 1], p.epsilon, p.pi);
                c.aim_id = n - 1;
            }
        }
    }
    const P3 ego{h.x, h.y, h.dir};
    const P3 aim{c.aim_x, c.aim_y, c.aim_dir};
    P2 road_points[DP_PATH_POINTS];
    std::memset(road_points, 0, sizeof(road_points));
    P2 last[DP_PATH_POINTS];

    // ---- InitialPlanning on the first cycle (Planning.cpp:124-128, 596-611) ----
    if (c.plan_count == 0) {
        spec::bezier_planning(ego, aim, road_points, DP_PATH_POINTS);
        for (int i = 0; i < DP_PATH_POINTS; ++i) { st.last_x[i] = road_points[i].x; st.last_y[i] = road_points[i].y; }
    }
    for (int i = 0; i < DP_PATH_POINTS; ++i) last[i] = P2{st.last_x[i], st.last_y[i]};

    // ---- GetVhclLocalState (Planning.cpp:623-676) ----
    double mind = 9999;
    int near_id = c.path_near_id;                       // member keeps its value if no point is closer than 9999
    for (int i = 0; i < DP_PATH_POINTS; ++i) {
        double d = std::sqrt((h.x - last[i].x) * (h.x - last[i].x) + (h.y - last[i].y) * (h.y - last[i].y));
        if (d < mind) { mind = d; near_id = i; }
    }
    const int front_id = near_id + 8;
    int idx = (near_id == 199) ? near_id - 1 : near_id;
    if (idx < 0 || idx > 198) { ++k.out.ub_hits; idx = idx < 0 ? 0 : 198; }
    const P2 pt = last[idx], pt_next = last[idx + 1];
    const double lat = get_lat_dis(P2{h.x, h.y}, pt, pt_next, p.epsilon);
    double remain = 0;
    for (int i = front_id; i < 199; ++i) {
        if (i < 0) { ++k.out.ub_hits; continue; }
        remain += std::sqrt((last[i + 1].x - last[i].x) * (last[i + 1].x - last[i].x) + (last[i + 1].y - last[i].y) * (last[i + 1].y - last[i].y));
    }
    const double dir_err = get_angle_err(get_road_angle(pt, pt_next, p.epsilon, p.pi), h.dir);
    c.path_near_id = near_id;

    // ---- UpdatePlanJudge (Planning.cpp:797-832) ----
    int cause = 0;
    bool afresh = true;
    if (c.plan_his_behavior != d_behavior) cause = 1;
    else if (std::fabs(lat) > 0.2) cause = 2;
    else if (std::fabs(dir_err) > 45) cause = 3;
    else if (pos == 0 && remain < p.road_remain_distance) cause = 4;
    else if (pos != 0 && remain < p.inter_remain_distance) cause = 4;
    else afresh = false;

    // ---- PathPlanning (Planning.cpp:845-877) or reuse (:142-146) ----
    if (afresh) {
        std::memset(road_points, 0, sizeof(road_points));
        if (pos == 0) spec::bezier_planning(ego, aim, road_points, DP_PATH_POINTS);
        else if (pos == 1 || pos == 2) {
            P2 refp[DP_PATH_POINTS];
            std::memset(refp, 0, sizeof(refp));
            int n = c.aim_id;
            if (n > DP_PATH_POINTS) { ++k.out.ub_hits; n = DP_PATH_POINTS; }     // reference overflows Ref_points[200] here
            for (int i = 0; i < n; ++i) {
                if (i >= (int)refpath.size()) { ++k.out.ub_hits; break; }
                refp[i] = refpath[i];
            }
            spec::mean_points(refp, n, road_points, DP_PATH_POINTS);
        }
    } else {
        std::memcpy(road_points, last, sizeof(last));
    }

    // ---- local path collision (Planning.cpp:152-168) ----
    Path rem;
    for (int i = near_id; i < DP_PATH_POINTS; ++i) {
        if (i < 0) { ++k.out.ub_hits; continue; }
        rem.push_back(road_points[i]);
    }
    spec::SearchResult s = search(k, rem, (double)(float)(-1.1), (double)(float)(1.1), k.out.trace ? &k.out.trace->local : nullptr);

    // ---- SpeedPlanning (Planning.cpp:888-990): identical for pos 0/1/2 ----
    double brake = 0, des_acc = 0;
    bool acc_flag = false;
    if (pos <= 2) {
        if (s.found) {
            if (s.dis_lng - 4 > 9) { brake = 3 + (s.dis_lng - 9) / (faraim - 9) * (v_exp - 3); }
            else if (s.dis_lng - 4 > 5) brake = 3;
            else { brake = 0; acc_flag = true; des_acc = -3; }
        } else brake = v_exp;
    }

    // ---- CalculateRadius on the PREVIOUS path (Planning.cpp:199, 1000-1019) ----
    auto lp = [&](int i) { if (i < 0 || i > 199) { ++k.out.ub_hits; i = i < 0 ? 0 : 199; } return last[i]; };
    const int mid = (near_id + front_id) / 2;
    const P2 a = lp(near_id), b = lp(mid), f = lp(front_id);
    const double d1 = std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y));
    const double d2 = std::sqrt((b.x - f.x) * (b.x - f.x) + (b.y - f.y) * (b.y - f.y));
    const double d3 = std::sqrt((a.x - f.x) * (a.x - f.x) + (a.y - f.y) * (a.y - f.y));
    const double dd = d1 * d1 + d2 * d2 - d3 * d3;
    const double cosA = dd / (2 * d1 * d2);
    const double sinA = std::sqrt(1 - cosA * cosA);
    const double radius = (sinA < 0.001) ? 1000 : 0.5 * d3 / sinA;

    // ---- outputs (Planning.cpp:173-214) and history (:216-223) ----
    r.path_lat_dis = lat; r.path_dir_err = dir_err; r.remain_dis = remain;
    r.mindist_lat = s.dis_lat; r.mindist_lon = s.dis_lng;
    r.brakespeed = brake; r.des_acc = des_acc; r.radius = radius;
    r.aim_x = c.aim_x; r.aim_y = c.aim_y; r.aim_dir = c.aim_dir; r.aim_id = c.aim_id;
    r.afresh_cause = (uint16_t)cause; r.afresh_planning = afresh;
    r.path_near_id = (int16_t)near_id; r.path_front_near_id = (int16_t)front_id;
    r.ob_index = (int16_t)s.ob_index; r.ob_pathid = (uint16_t)s.pathid;
    r.ob_flag = s.found; r.acc_flag = acc_flag;
    r.cnt = (uint8_t)(c.plan_count % 100);
    if (k.out.path_xy)
        for (int i = 0; i < DP_PATH_POINTS; ++i) { k.out.path_xy[i] = road_points[i].x; k.out.path_xy[DP_PATH_POINTS + i] = road_points[i].y; }
    if (k.out.path_ll) {
        spec::Datum dm{p.lat0, p.lng0, p.k_lat, p.k_lng};
        for (int i = 0; i < DP_OUT_POINTS; ++i)
            spec::global_to_wgs84(dm, road_points[2 * i].x, road_points[2 * i].y, &k.out.path_ll[i], &k.out.path_ll[DP_OUT_POINTS + i]);
    }
    c.plan_his_behavior = d_behavior;
    for (int i = 0; i < DP_PATH_POINTS; ++i) { st.last_x[i] = road_points[i].x; st.last_y[i] = road_points[i].y; }
    uint8_t cnt = (uint8_t)(c.plan_count + 1);
    if (cnt % 100 == 1) cnt = 1;
    c.plan_count = cnt;
}

}  // namespace

void cycle(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const double* ox, const double* oy,
           SceneState& st, CycleOut& out, bool exhaustive_sweep) {
    std::vector<P2> obs(h.n_obs);
    for (int i = 0; i < h.n_obs; ++i) obs[i] = P2{ox[i], oy[i]};
    out.n_calls = 0;
    out.ub_hits = 0;
    std::memset(out.rec, 0, sizeof(*out.rec));
    if (out.trace) std::memset(out.trace, 0, sizeof(*out.trace));
    Ctx k{m, p, h, obs.data(), (int)obs.size(), out, exhaustive_sweep};
    Path refpath;
    dp_carry& c = st.c;
    // ---- Decision thread iteration (Decision.cpp:171-201) ----
    if (h.pos == 0) segment_decision(k, st, refpath);
    else if (h.pos == 1 || h.pos == 2) junction_decision(k, st, refpath);
    dp_plan_record& r = *out.rec;
    r.behavior = c.behavior; r.target_roadnum = c.target_roadnum; r.target_lanenum = c.target_lanenum;
    r.light = c.light_status; r.velocity_expect = c.velocity_expect; r.behavior_to_dlg = c.behavior_to_dlg;
    c.his_behavior = c.behavior; c.his_light_status = c.light_status; c.his_target_lanenum = c.target_lanenum;
    if (out.trace) out.trace->refpath_len = (uint16_t)refpath.size();
    // ---- Planning thread iteration ----
    planning_cycle(k, st, refpath);
    r.n_traj = (uint16_t)k.n_traj;
    if (out.trace) { out.trace->ub_hits = (uint16_t)out.ub_hits; out.trace->pts_scored = (uint32_t)k.pts; }
}

}  // namespace oracle
