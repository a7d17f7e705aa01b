// csrc/dp_cycle.cu -- the fused Decision + Planning cycle kernel (sm_100a).
//
// One WARP owns one scene for the whole cycle: LoadRefPath -> AroundObstacle -> BehaviorDecision
// (with the in-lane avoid sweep) -> SpeedDecision/RefPath  (Decision.cpp:216-315, 323-486), then
// Calculate_aim_dis -> SearchAimPoint -> GetVhclLocalState -> UpdatePlanJudge -> PathPlanning ->
// local-path SearchObstacle -> SpeedPlanning -> CalculateRadius -> result packing
// (Planning.cpp:118-217).  Scalar rule-tree state is warp-uniform (every lane holds it, branches
// never diverge); the geometry operators in dp_device.cuh spread their inner loops over the lanes.
// Candidate paths exist only as shared-memory tiles; HBM sees the 128-byte scene header, the
// obstacle SoA rows, the 128-byte carry, the previous local path (the reference's own cross-cycle
// state, Planning.cpp:6) and the 128-byte plan record.
#include "dp_device.cuh"
#include "dp_kernels.h"

namespace {

struct Beh { int behavior, target, light; bool lanechg, obsavoid; int dlg; };

__device__ __forceinline__ void put_slot(dp_search_slot* slot, const SearchRes& r, int evaluated, int lane) {
    if (slot && lane == 0) {
        slot->dis_lat = r.dis_lat; slot->dis_lng = r.dis_lng; slot->ob_index = (int16_t)r.ob;
        slot->pathid = (uint16_t)r.pathid; slot->evaluated = (uint8_t)evaluated; slot->found = r.found ? 1 : 0;
        slot->pad[0] = slot->pad[1] = 0;
    }
}

// forward 120-point / backward 40-point slices of one lane (Decision.cpp:581-596, 611-622, 649-660)
__device__ __forceinline__ void lane_paths(const DevMap& m, int gl, int id, int id_more, PathSrc& fwd, PathSrc& rear, int& ub) {
    const int off = m.lane_pt_off[gl], n = m.lane_pt_off[gl + 1] - off;
    const int a = min(n, id + id_more), b = min(n, id + 120 + id_more);
    fwd.kind = 0; fwd.base0 = off + a; fwd.step0 = 1; fwd.n0 = max(0, b - a); fwd.P = fwd.n0; fwd.base1 = 0; fwd.d = 0.0; fwd.s0 = 0;
    const int lo = max(0, id + id_more - 40);
    int start = a;
    if (start >= n) { start = n - 1; if (a > lo) ++ub; }   // reference reads index == size here (UB); clamp
    rear.kind = 0; rear.base0 = off + start; rear.step0 = -1; rear.n0 = max(0, a - lo); rear.P = rear.n0; rear.base1 = 0; rear.d = 0.0; rear.s0 = 0;
}

// does the arclength from `id` along lane `gl`, while cond(attr[i+1]) holds, exceed thr?
// (Decision.cpp:1179-1190 and siblings; partial sums are monotone, so stopping at the first
// partial sum > thr gives the reference's answer without walking the whole lane)
// mode: 0 = attr == 1, 1 = attr & 1, 2 = never
__device__ bool run_exceeds(const DevMap& m, WarpSmem& sm, int gl, int id, int mode, double thr, int lane) {
    if (mode == 2) return 0.0 > thr;
    const int off = m.lane_pt_off[gl], n = m.lane_pt_off[gl + 1] - off;
    double sum = 0.0;
    for (int i0 = id; i0 < n - 1; i0 += DP_SCR) {
        const int cnt = min(DP_SCR, n - 1 - i0);
        int first_bad = cnt;                               // first term whose condition fails
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            bool ok = true;
            if (j < cnt) {
                const int a = m.attr[off + i0 + j + 1];
                ok = (mode == 0) ? (a == 1) : ((a & 1) != 0);
                sm.scr[j] = m.lenp[off + i0 + j];
            }
            const unsigned bad = __ballot_sync(DP_FULL, !ok);
            if (bad) { first_bad = j0 + __ffs(bad) - 1; break; }
        }
        __syncwarp();
        const int cn = min(cnt, first_bad);
        bool over = false;
        for (int j = 0; j < cn; ++j) { sum += sm.scr[j]; if (sum > thr) { over = true; break; } }
        __syncwarp();
        if (over) return true;
        if (first_bad < cnt) break;
    }
    return sum > thr;
}

// SearchAimPoint along one map lane (Planning.cpp:410-432 and the lane-change twins)
__device__ void aim_walk_lane(const DevMap& m, WarpSmem& sm, int gl_walk, int from, int to_excl, int gl_fb, int fb_idx, int fb_id,
                              float faraim, dp_carry& c, int lane, int& ub) {
    if (from >= to_excl) return;
    const int off = m.lane_pt_off[gl_walk], n = m.lane_pt_off[gl_walk + 1] - off;
    if (from < 0 || to_excl > n - 1) { ++ub; if (from < 0) return; to_excl = min(to_excl, n - 1); if (from >= to_excl) return; }
    double sum = 0.0;
    int hit = -1;
    for (int i0 = from; i0 < to_excl && hit < 0; i0 += DP_SCR) {
        const int cnt = min(DP_SCR, to_excl - i0);
        for (int j = lane; j < cnt; j += 32) sm.scr[j] = m.lenp[off + i0 + j];
        __syncwarp();
        for (int j = 0; j < cnt; ++j) { sum += sm.scr[j]; if (sum - 4 > faraim) { hit = i0 + j; break; } }
        __syncwarp();
    }
    if (hit >= 0) {
        c.aim_x = m.x[off + hit]; c.aim_y = m.y[off + hit]; c.aim_dir = m.dir[off + hit]; c.aim_id = hit;
    } else {
        const int offb = m.lane_pt_off[gl_fb], nb = m.lane_pt_off[gl_fb + 1] - offb;
        int fi = fb_idx;
        if (fi < 0 || fi >= nb) { ++ub; fi = fi < 0 ? 0 : nb - 1; }
        c.aim_x = m.x[offb + fi]; c.aim_y = m.y[offb + fi]; c.aim_dir = m.dir[offb + fi]; c.aim_id = fb_id;
    }
}

}  // namespace

__global__ void __launch_bounds__(DP_WARPS_PER_BLOCK * 32)
dp_cycle_kernel(DevMap m, dp_params p, int n_scenes, const dp_scene_hdr* __restrict__ hdr, const double* __restrict__ obs_x,
                const double* __restrict__ obs_y, int max_obs, dp_carry* __restrict__ carry, double* __restrict__ last_path,
                dp_plan_record* __restrict__ rec, dp_trace_record* __restrict__ trace, double* __restrict__ path_xy,
                double* __restrict__ path_ll) {
    __shared__ WarpSmem smem[DP_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int scene = blockIdx.x * DP_WARPS_PER_BLOCK + wib;
    if (scene >= n_scenes) return;
    WarpSmem& sm = smem[wib];
    const dp_scene_hdr& h = hdr[scene];
    const double* ox = obs_x + (size_t)scene * max_obs;
    const double* oy = obs_y + (size_t)scene * max_obs;
    const int N = min((int)h.n_obs, max_obs);
    dp_carry c = carry[scene];
    double* lastp = last_path + (size_t)scene * 2 * DP_PATH_POINTS;
    dp_trace_record* tr = trace ? trace + scene : nullptr;
    if (tr) {                                               // zero the trace record cooperatively
        uint32_t* w = reinterpret_cast<uint32_t*>(tr);
        for (int i = lane; i < (int)(sizeof(dp_trace_record) / 4); i += 32) w[i] = 0;
        __syncwarp();
    }
    const double Vw = p.vehicle_width;
    const int pos = h.pos;
    int ub = 0, n_traj = 0, sweep_pick = -1, pts = 0;
    PathSrc refpath;                                        // DecisionOut.refpath as a recipe
    refpath.kind = 0; refpath.P = 0; refpath.base0 = 0; refpath.step0 = 1; refpath.n0 = 0; refpath.base1 = 0; refpath.s0 = 0; refpath.d = 0.0;
    refpath.gx = nullptr; refpath.gy = nullptr;

    if (pos == 0) {
        // =========================== SegmentDecision (Decision.cpp:216-315) ===========================
        const int road = h.road_num, lane_n = h.lane_num;
        const int gl = m.road_lane_base[road - 1] + lane_n - 1;
        const int lane_sum = m.road_lane_base[road] - m.road_lane_base[road - 1];
        const int off = m.lane_pt_off[gl], id_sum = m.lane_pt_off[gl + 1] - off;
        const int id = (int)(uint16_t)h.id[lane_n - 1];
        // ---- Nav_LaneChange (Decision.cpp:685-738) ----
        unsigned navi = 4, navi_t = 0;
        for (int i = 0; i < DP_LANESUM && h.out_lane_no[i] != 0; ++i)
            if (lane_n == h.out_lane_no[i]) { navi = 0; break; }
        int out_min = h.out_lane_no[0], out_max = 1;
        for (int i = 0; i < DP_LANESUM; ++i) if (h.out_lane_no[i] > out_max) out_max = h.out_lane_no[i];
        if (navi == 4) {
            int dirn = 0;
            if (lane_n < out_min) { navi = 2; dirn = 2; }
            else if (lane_n > out_max) { navi = 1; dirn = 1; }
            else navi = 0;
            if (dirn) {                                     // CalcNaviLaneChgTimes (Decision.cpp:498-538)
                int times = 5;
                for (int i = 0; i < DP_LANESUM && h.out_lane_no[i] != 0; ++i) {
                    const int t = (dirn == 1) ? lane_n - h.out_lane_no[i] : h.out_lane_no[i] - lane_n;
                    if (t < times) times = t;
                }
                navi_t = (unsigned)(times & 0xff);
            }
        }
        // ---- LoadRefPath (Decision.cpp:553-673) ----
        const int idc = min(id, id_sum - 1);
        const int lanechg = m.attr[off + idc];
        const double W = m.width[off + idc] / 100.0;
        PathSrc reg[6];                                     // F, R, LF, LR, RF, RR
#pragma unroll
        for (int r = 0; r < 6; ++r) { reg[r].kind = 0; reg[r].P = 0; reg[r].n0 = 0; reg[r].base0 = 0; reg[r].step0 = 1; reg[r].base1 = 0; reg[r].s0 = 0; reg[r].d = 0.0; reg[r].gx = nullptr; reg[r].gy = nullptr; }
        lane_paths(m, gl, id, p.id_more, reg[0], reg[1], ub);
        if (lanechg == 1 || lanechg == 3) {
            if (lane_n > 1) {
                const int idl = (int)(uint16_t)h.id[lane_n - 2];
                const int nl = m.lane_pt_off[gl] - m.lane_pt_off[gl - 1];
                if (idl > 0 && idl < nl) lane_paths(m, gl - 1, idl, p.id_more, reg[2], reg[3], ub);
            } else {
                reg[2] = reg[0]; reg[2].d = -1 * W;
                reg[3] = reg[1]; reg[3].d = -1 * W;
            }
        }
        if (lanechg == 2) {
            if (lane_n < lane_sum) {
                const int idr = (int)(uint16_t)h.id[lane_n];
                const int nr = m.lane_pt_off[gl + 2] - m.lane_pt_off[gl + 1];
                if (idr > 0 && idr < nr) lane_paths(m, gl + 1, idr, p.id_more, reg[4], reg[5], ub);
            } else {
                reg[4] = reg[0]; reg[4].d = W;
                reg[5] = reg[1]; reg[5].d = W;
            }
        }
        // ---- AroundObstacle (Decision.cpp:759-881) ----
        double gap[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            gap[r] = 0.0;                                   // memset state when the path is empty
            if (reg[r].P != 0) {
                const double lo = (r < 4) ? -0.5 * Vw : -0.5 * W;
                const double hi = (r < 2) ? 0.5 * Vw : (r < 4 ? 0.5 * W : 0.5 * Vw);
                const SearchRes s = dp_search_path(m, sm, reg[r], ox, oy, N, lo, hi, lane);
                gap[r] = s.dis_lng;
                ++n_traj; pts += reg[r].P;
                put_slot(tr ? &tr->region[r] : nullptr, s, 1, lane);
            }
        }
        const double gF = gap[0], gLF = gap[2], gLR = gap[3], gRF = gap[4], gRR = gap[5];
        if (tr && lane == 0) { tr->width_curlane = W; tr->navi_lanechg = navi; tr->navi_lanechg_times = navi_t; }

        // ---- BehaviorDecision (Decision.cpp:898-1773) ----
        Beh cur;
        cur.behavior = c.behavior; cur.target = c.target_lanenum; cur.light = c.light_status;
        cur.lanechg = c.lanechg_status != 0; cur.obsavoid = c.obsavoid_status != 0; cur.dlg = c.behavior_to_dlg;
        const int his_behavior = c.his_behavior, his_target = c.his_target_lanenum, his_light = c.his_light_status;
        int lane_cur = lane_n;
        int z_light = c.light_status;
        const double period = h.period_ms;
        int K = 0;
        while (K < DP_MAX_SWEEP && (double)K < (W - Vw) / 0.6) ++K;
#define DP_KEEP(reset) do { cur.behavior = 1; cur.target = lane_cur; if (reset) cur.lanechg = false; } while (0)
        if (lanechg == 0) {                                 // :920-1010
            if (gF < 15) {
                c.no_obsavoid_time = 0;
                c.obsavoid_time++;
                if (c.obsavoid_time > 2) {
                    // in-lane avoid sweep: candidates L0..L(K-1), R0..R(K-1), first feasible wins
                    // (Decision.cpp:940-974).  With a trace buffer the remaining candidates are
                    // scored too (evaluated = 2) but neither counted nor used.
                    bool picked = false;
                    for (int side = 0; side < 2; ++side) {
                        for (int i = 0; i < K; ++i) {
                            if (picked && !tr) break;
                            PathSrc cand = reg[0];
                            cand.d = (side == 0 ? -0.3 : 0.3) * i;
                            const SearchRes s = dp_search_path(m, sm, cand, ox, oy, N, -0.5 * Vw, 0.5 * Vw, lane);
                            put_slot(tr ? &tr->sweep[side * DP_MAX_SWEEP + i] : nullptr, s, picked ? 2 : 1, lane);
                            if (picked) continue;
                            ++n_traj; pts += cand.P;
                            if (s.dis_lng > 25) {
                                cur.behavior = side == 0 ? 4 : 5; cur.target = lane_cur; cur.light = side == 0 ? 1 : 2;
                                cur.obsavoid = true; cur.dlg = side == 0 ? 11 : 12;
                                picked = true; sweep_pick = side * K + i;
                            }
                        }
                    }
                } else { cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1; }
            } else {
                if (c.obsavoid_status == 0) { cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1; }
                else {
                    c.no_obsavoid_time++;
                    if (c.no_obsavoid_time > 3) { cur.behavior = 1; cur.target = lane_cur; cur.light = 0; cur.dlg = 1; cur.obsavoid = false; }
                }
                cur.dlg = 1;
            }
        } else {                                            // :1012-1772
            c.no_obsavoid_time = 0;
            c.obsavoid_time = 0;
            if (!cur.lanechg) {
                if (navi != 0) {                            // :1021-1144
                    if (navi == 1) {
                        if (lanechg == 1 || lanechg == 3) {
                            cur.dlg = 2;
                            if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                            c.leftlight_time += period;
                            if (((gLF > gF + 10) || (gLF > 40)) && gLR > 15 && c.leftlight_time > 2000) {
                                cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                            } else DP_KEEP(true);
                        } else { DP_KEEP(true); cur.dlg = 4; }
                    } else if (navi == 2) {
                        if (lanechg == 2 || lanechg == 3) {
                            cur.dlg = 3;
                            if (cur.light != 2) { cur.lanechg = true; c.rightlight_time = 0; }   // sic: lanechg_status = 2 (:1094)
                            c.rightlight_time += period;
                            if (((gRF > gF + 10) || (gRF > 40)) && gRR > 15 && c.rightlight_time >= 2000) {
                                cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                            } else DP_KEEP(true);
                        } else { DP_KEEP(true); cur.dlg = 4; }
                    }
                } else {                                    // :1146-1757
                    if (gF < (2 * 10 + 5)) {
                        c.frontobs_time++;
                        if (c.frontobs_time > 2) {
                            c.frontobs_time = 3;
                            bool has_m1 = false, has_p1 = false;    // exit-lane list contains lane-1 / lane+1
                            for (int i = 0; i < DP_LANESUM && h.out_lane_no[i] != 0; ++i) {
                                if (h.out_lane_no[i] == lane_cur - 1) has_m1 = true;
                                if (h.out_lane_no[i] == lane_cur + 1) has_p1 = true;
                            }
                            if (lanechg == 1) {             // :1157-1294
                                if (lane_cur > 1) {
                                    cur.dlg = 5;
                                    const bool no_back = !has_m1;
                                    bool chg = false;
                                    if (no_back) {
                                        if (run_exceeds(m, sm, gl, id, 0, 60.0, lane)) {
                                            chg = true;
                                            if (z_light != 1) { z_light = 1; c.leftlight_time = 0; }
                                            c.leftlight_time += period;
                                            if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                        } else z_light = 0;
                                    } else {
                                        if (run_exceeds(m, sm, gl, id, 1, 15.0, lane)) {
                                            chg = true;
                                            if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                                            c.leftlight_time += period;
                                            if (c.leftlight_time > 2000) c.leftlight_time = 2100;
                                        } else cur.light = 0;
                                    }
                                    if (chg && gLF > gF + 10 && gLR > 10 && c.leftlight_time > 1500) {
                                        c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                    } else DP_KEEP(true);
                                } else DP_KEEP(true);
                            } else if (lanechg == 2) {      // :1296-1424
                                if (lane_cur < lane_sum) {
                                    cur.dlg = 6;
                                    const bool no_back = !has_m1;                     // sic (:1307)
                                    bool chg = false;
                                    if (no_back) {
                                        if (run_exceeds(m, sm, gl, id, 1, 50.0, lane)) {
                                            chg = true;
                                            if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                            c.leftlight_time += period;
                                            if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                        }
                                    } else {
                                        if (run_exceeds(m, sm, gl, id, 1, 10.0, lane)) {
                                            chg = true;
                                            if (cur.light != 1) { cur.light = 1; c.leftlight_time = 0; }
                                            c.leftlight_time += period;
                                            if (c.leftlight_time > 2000) c.leftlight_time = 2000;
                                        }
                                    }
                                    if (chg && gRF > gF + 10 && gRR > 10 && c.leftlight_time > 1500) {
                                        c.frontobs_time = 0; cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                                    } else DP_KEEP(true);
                                } else DP_KEEP(true);
                            } else if (lanechg == 3) {      // :1426-1738
                                const bool nb_left = !has_m1, nb_right = !has_p1;
                                bool left_ok = false, right_ok = false;
                                if (lane_cur > 1) left_ok = nb_left ? run_exceeds(m, sm, gl, id, 1, 50.0, lane) : run_exceeds(m, sm, gl, id, 2, 10.0, lane);
                                if (lane_cur < lane_sum) right_ok = run_exceeds(m, sm, gl, id, 1, nb_right ? 50.0 : 10.0, lane);
#define DP_TICK do { c.leftlight_time += period; if (c.leftlight_time > 2000) c.leftlight_time = 2000; } while (0)
                                if (left_ok && !nb_left) {                          // :1545-1594
                                    if (!cur.lanechg) { cur.light = 1; c.leftlight_time = 0; }
                                    DP_TICK;
                                    if (gLF > gF + 10 && gLR > 10 && c.leftlight_time > 2000) {
                                        c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                    } else DP_KEEP(true);
                                } else if (right_ok && !nb_right) {                 // :1596-1636
                                    if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                    DP_TICK;
                                    if (gRF > gF + 10) {
                                        if (gRR > 10 && c.leftlight_time > 2000) {
                                            c.frontobs_time = 0; cur.behavior = 3; cur.target = lane_cur + 1; cur.lanechg = true;
                                        } else DP_KEEP(false);
                                    }
                                } else if (left_ok) {                               // :1638-1686
                                    if (cur.light != 1) { cur.lanechg = true; c.leftlight_time = 0; }   // sic (:1642)
                                    DP_TICK;
                                    if (gLF > gF + 10 && gLR > 10 && c.leftlight_time > 2000) {
                                        c.frontobs_time = 0; cur.behavior = 2; cur.target = lane_cur - 1; cur.lanechg = true;
                                    } else DP_KEEP(false);
                                } else if (right_ok) {                              // :1688-1730
                                    if (cur.light != 2) { cur.light = 2; c.leftlight_time = 0; }
                                    DP_TICK;
                                    if (gRF > gF + 10) {
                                        if (gRR > 10 && c.leftlight_time > 2000) {
                                            c.frontobs_time = 0; cur.behavior = 3; lane_cur = 1; cur.target = 1; cur.lanechg = true;   // sic (:1712)
                                        } else DP_KEEP(false);
                                    }
                                } else DP_KEEP(true);
                            }
                        } else DP_KEEP(true);               // :1741-1746
                    } else { c.frontobs_time = 0; cur.dlg = 8; DP_KEEP(true); }
                }
            } else {                                        // lane change in progress, :1760-1771
                cur.dlg = 9;
                if (cur.target == lane_cur) { cur.lanechg = false; cur.light = 0; }
                cur.behavior = his_behavior; cur.target = his_target; cur.light = his_light;
            }
        }
        (void)z_light;
        // ---- SpeedDecision / RefPath / write-back (Decision.cpp:1781-1816, 307-313) ----
        c.velocity_expect = (cur.behavior == 4 || cur.behavior == 5) ? 5 : 10;
        refpath = (cur.behavior == 2) ? reg[2] : (cur.behavior == 3) ? reg[4] : reg[0];
        c.behavior = (uint16_t)cur.behavior; c.light_status = (uint16_t)cur.light; c.target_lanenum = (uint16_t)cur.target;
        c.lanechg_status = cur.lanechg; c.obsavoid_status = cur.obsavoid; c.behavior_to_dlg = (uint16_t)cur.dlg;
        c.target_roadnum = h.road_num;
    } else if (pos == 1 || pos == 2) {
        // =================== PreStubDecision / StubDecision (Decision.cpp:323-486) ===================
        const int gc = (h.conn >= 0 && h.conn < m.n_conn) ? m.conn[h.conn].lane : -1;
        const int coff = gc >= 0 ? m.lane_pt_off[gc] : 0, n_inter = gc >= 0 ? m.lane_pt_off[gc + 1] - coff : 0;
        const int gl = m.road_lane_base[h.road_num - 1] + h.lane_num - 1;
        const int off = m.lane_pt_off[gl], n = m.lane_pt_off[gl + 1] - off;
        if (pos == 1) {
            const int id = (int)(uint16_t)h.id[h.lane_num - 1];
            refpath.base0 = off + id; refpath.step0 = 1; refpath.n0 = max(0, n - id);
            refpath.base1 = coff; refpath.P = refpath.n0 + n_inter;
        } else {
            const int id = (int)(uint16_t)h.id[h.last_lanenum - 1];
            refpath.base0 = coff + id; refpath.step0 = 1; refpath.n0 = max(0, n_inter - id);
            refpath.base1 = off; refpath.P = refpath.n0 + min(60, n);
        }
        const SearchRes s = dp_search_path(m, sm, refpath, ox, oy, N, -0.5 * Vw, 0.5 * Vw, lane);
        ++n_traj; pts += refpath.P;
        put_slot(tr ? &tr->junction : nullptr, s, 1, lane);
        if (s.dis_lng < 13) {
            const double v = s.dis_lng - 3;
            c.velocity_expect = v > 0 ? v : 0;
            c.behavior_to_dlg = 13;
        } else { c.velocity_expect = 10; c.behavior_to_dlg = 1; }
        c.light_status = (h.stub_attribute == 3) ? 1 : h.stub_attribute;
        c.behavior = 1;
        c.target_roadnum = h.road_num;
        c.target_lanenum = h.lane_num;
    }
    // ---- tail of the Decision thread iteration (Decision.cpp:187-201) ----
    const int d_behavior = c.behavior, d_target = c.target_lanenum;
    const double v_exp = c.velocity_expect;
    c.his_behavior = c.behavior; c.his_light_status = c.light_status; c.his_target_lanenum = c.target_lanenum;
    if (tr && lane == 0) tr->refpath_len = (uint16_t)refpath.P;

    // ================================ Planning thread iteration ================================
    // ---- Calculate_aim_dis (Planning.cpp:242-290), FLOAT faraim_dis ----
    float faraim = 0.f;
    if (pos == 0) {
        faraim = (float)((h.velocity / 3.6) * 5 + 4);
        if (faraim > p.road_faraim_max) faraim = (float)p.road_faraim_max;
        else if (faraim < p.road_faraim_min) faraim = (float)p.road_faraim_min;
    } else if (pos == 1) faraim = (float)p.pre_inter_faraim;
    else if (pos == 2) faraim = (float)p.inter_faraim;
    if (tr && lane == 0) tr->faraim_dis = faraim;

    // ---- SearchAimPoint (Planning.cpp:303-583) ----
    if (pos == 0) {
        const int road = h.road_num, lane_n = h.lane_num;
        const int gl = m.road_lane_base[road - 1] + lane_n - 1;
        const int lane_sum = m.road_lane_base[road] - m.road_lane_base[road - 1];
        const int cur_id = h.id[lane_n - 1], cur_sum = m.lane_pt_off[gl + 1] - m.lane_pt_off[gl];
        int left_id = 0, left_sum = 0, right_id = 0, right_sum = 0;
        if (lane_n > 1) { left_id = h.id[lane_n - 2]; left_sum = m.lane_pt_off[gl] - m.lane_pt_off[gl - 1]; }
        if (lane_n < lane_sum) { right_id = h.id[lane_n]; right_sum = m.lane_pt_off[gl + 2] - m.lane_pt_off[gl + 1]; }
        if (d_target == lane_n) {
            if (d_behavior == 1) aim_walk_lane(m, sm, gl, cur_id, cur_sum - 1, gl, cur_sum - 1, cur_sum - 1, faraim, c, lane, ub);
        } else {
            if (d_behavior == 2) {
                if (lane_n > 1) aim_walk_lane(m, sm, gl - 1, left_id, left_sum - 1, gl, left_sum - 2, left_sum - 1, faraim, c, lane, ub);
            } else if (d_behavior == 3) {
                if (lane_n < lane_sum) aim_walk_lane(m, sm, gl + 1, right_id, left_sum - 1, gl + 1, right_sum - 1, right_sum - 1, faraim, c, lane, ub);
                else if (right_id < left_sum - 1) ++ub;
            }
        }
    } else if (pos == 1 || pos == 2) {
        const int n = refpath.P;
        double sum = 0.0;
        int hit = -1;
        for (int i0 = 0; i0 < n - 1 && hit < 0; i0 += DP_SCR) {
            const int cnt = min(DP_SCR, n - 1 - i0);
            for (int j = lane; j < cnt; j += 32) {
                const double2 a = dp_path_point(m, sm, refpath, i0 + j), b = dp_path_point(m, sm, refpath, i0 + j + 1);
                sm.scr[j] = dp_dist_plain(a.x, a.y, b.x, b.y);
            }
            __syncwarp();
            for (int j = 0; j < cnt; ++j) { sum += sm.scr[j]; if ((sum - 4) > faraim) { hit = i0 + j; break; } }
            __syncwarp();
        }
        if (hit >= 0) {
            const double2 a = dp_path_point(m, sm, refpath, hit);
            c.aim_x = a.x; c.aim_y = a.y;
            if ((size_t)hit < (size_t)n - 4) {
                const double2 b = dp_path_point(m, sm, refpath, hit + 2);
                c.aim_dir = dp_heading(a.x, a.y, b.x, b.y, p.epsilon, p.pi);
            } else {
                int ia = hit - 2; if (ia < 0) { ++ub; ia = 0; }
                const double2 b = dp_path_point(m, sm, refpath, ia);
                c.aim_dir = dp_heading(b.x, b.y, a.x, a.y, p.epsilon, p.pi);
            }
            c.aim_id = hit;
        } else if (n - 1 > 0) {
            int ia = n - 3; if (ia < 0) { ++ub; ia = 0; }
            const double2 e = dp_path_point(m, sm, refpath, n - 1), b = dp_path_point(m, sm, refpath, ia);
            c.aim_x = e.x; c.aim_y = e.y;
            c.aim_dir = dp_heading(b.x, b.y, e.x, e.y, p.epsilon, p.pi);
            c.aim_id = n - 1;
        }
    }

    // ---- InitialPlanning on the first cycle (Planning.cpp:124-128) ----
    if (c.plan_count == 0) {
        dp_bezier_to_plan(sm, h.x, h.y, h.dir, c.aim_x, c.aim_y, c.aim_dir, lane);
        for (int i = lane; i < DP_PATH_POINTS; i += 32) { lastp[i] = sm.plan[i].x; lastp[DP_PATH_POINTS + i] = sm.plan[i].y; }
        __syncwarp();
    }
    // ---- GetVhclLocalState (Planning.cpp:623-676) on last_Bpoints ----
    double md = 9999.0;
    int mi = 0x7fffffff;
    for (int i = lane; i < DP_PATH_POINTS; i += 32) {
        const double d = dp_dist_plain(h.x, h.y, lastp[i], lastp[DP_PATH_POINTS + i]);
        if (d < md) { md = d; mi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(DP_FULL, md, o);
        const int oi = __shfl_xor_sync(DP_FULL, mi, o);
        if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
    }
    const int near_id = (mi == 0x7fffffff) ? c.path_near_id : mi;
    const int front_id = near_id + 8;
    int idx = (near_id == 199) ? near_id - 1 : near_id;
    if (idx < 0 || idx > 198) { ++ub; idx = idx < 0 ? 0 : 198; }
    const double ptx = lastp[idx], pty = lastp[DP_PATH_POINTS + idx], pnx = lastp[idx + 1], pny = lastp[DP_PATH_POINTS + idx + 1];
    const double lat = dp_lat_dis(h.x, h.y, ptx, pty, pnx, pny, p.epsilon);
    double remain = 0.0;
    {
        const int f0 = max(front_id, 0);
        if (front_id < 0) ++ub;
        const int cnt = max(0, 199 - f0);
        for (int j = lane; j < cnt; j += 32) {
            const int i = f0 + j;
            sm.scr[j] = dp_dist_plain(lastp[i + 1], lastp[DP_PATH_POINTS + i + 1], lastp[i], lastp[DP_PATH_POINTS + i]);
        }
        __syncwarp();
        remain = dp_seq_sum(sm, cnt, 0.0);
        __syncwarp();
    }
    const double dir_err = dp_angle_err(dp_heading(ptx, pty, pnx, pny, p.epsilon, p.pi), h.dir);
    c.path_near_id = near_id;

    // ---- UpdatePlanJudge (Planning.cpp:797-832) ----
    int cause = 0;
    bool afresh = true;
    if (c.plan_his_behavior != d_behavior) cause = 1;
    else if (fabs(lat) > 0.2) cause = 2;
    else if (fabs(dir_err) > 45) cause = 3;
    else if (pos == 0 && remain < p.road_remain_distance) cause = 4;
    else if (pos != 0 && remain < p.inter_remain_distance) cause = 4;
    else afresh = false;

    // ---- CalculateRadius reads the PREVIOUS path (Planning.cpp:199, 1000-1019): do it before overwriting ----
    double radius;
    {
        int ia = near_id, ib = (near_id + front_id) / 2, ic = front_id;
        if (ia < 0 || ia > 199) { ++ub; ia = ia < 0 ? 0 : 199; }
        if (ib < 0 || ib > 199) { ++ub; ib = ib < 0 ? 0 : 199; }
        if (ic < 0 || ic > 199) { ++ub; ic = ic < 0 ? 0 : 199; }
        const double ax = lastp[ia], ay = lastp[DP_PATH_POINTS + ia], bx = lastp[ib], by = lastp[DP_PATH_POINTS + ib];
        const double fx = lastp[ic], fy = lastp[DP_PATH_POINTS + ic];
        const double d1 = dp_dist_plain(ax, ay, bx, by), d2 = dp_dist_plain(bx, by, fx, fy), d3 = dp_dist_plain(ax, ay, fx, fy);
        const double dd = d1 * d1 + d2 * d2 - d3 * d3;
        const double cosA = dd / (2 * d1 * d2);
        const double sinA = sqrt(1 - cosA * cosA);
        radius = (sinA < 0.001) ? 1000 : 0.5 * d3 / sinA;
    }

    // ---- PathPlanning (Planning.cpp:845-877) or reuse (:142-146): road_points -> sm.plan ----
    if (afresh) {
        if (pos == 0) dp_bezier_to_plan(sm, h.x, h.y, h.dir, c.aim_x, c.aim_y, c.aim_dir, lane);
        else if (pos == 1 || pos == 2) {
            int n = c.aim_id;
            if (n > DP_PATH_POINTS) { ++ub; n = DP_PATH_POINTS; }
            if (n > refpath.P) { ++ub; n = refpath.P; }
            dp_mean_points_to_plan(m, sm, refpath, n, lane);
        } else {
            for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = make_double2(0.0, 0.0);
            __syncwarp();
        }
    } else if (c.plan_count != 0) {
        for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = make_double2(lastp[i], lastp[DP_PATH_POINTS + i]);
        __syncwarp();
    }                                                       // (plan_count == 0 && !afresh: sm.plan already holds the initial path)

    // ---- local path collision (Planning.cpp:152-168) ----
    PathSrc rem;
    rem.kind = 1; rem.s0 = max(near_id, 0); rem.P = DP_PATH_POINTS - rem.s0; rem.d = 0.0;
    rem.base0 = rem.step0 = rem.n0 = rem.base1 = 0; rem.gx = rem.gy = nullptr;
    const SearchRes ls = dp_search_path(m, sm, rem, ox, oy, N, (double)(float)(-1.1), (double)(float)(1.1), lane);
    ++n_traj; pts += rem.P;
    put_slot(tr ? &tr->local : nullptr, ls, 1, lane);

    // ---- SpeedPlanning (Planning.cpp:888-990) ----
    double brake = 0.0, des_acc = 0.0;
    bool acc_flag = false;
    if (pos <= 2) {
        if (ls.found) {
            if (ls.dis_lng - 4 > 9) brake = 3 + (ls.dis_lng - 9) / (faraim - 9) * (v_exp - 3);
            else if (ls.dis_lng - 4 > 5) brake = 3;
            else { brake = 0; acc_flag = true; des_acc = -3; }
        } else brake = v_exp;
    }

    // ---- outputs (Planning.cpp:173-214) and history (:216-223) ----
    for (int i = lane; i < DP_PATH_POINTS; i += 32) {
        const double2 q = sm.plan[i];
        lastp[i] = q.x; lastp[DP_PATH_POINTS + i] = q.y;
        if (path_xy) { path_xy[(size_t)scene * 400 + i] = q.x; path_xy[(size_t)scene * 400 + DP_PATH_POINTS + i] = q.y; }
    }
    if (path_ll) {
        for (int i = lane; i < DP_OUT_POINTS; i += 32) {
            const double2 q = sm.plan[2 * i];
            path_ll[(size_t)scene * 200 + i] = fma(q.y, p.k_lat, p.lat0);
            path_ll[(size_t)scene * 200 + DP_OUT_POINTS + i] = fma(q.x, p.k_lng, p.lng0);
        }
    }
    const uint8_t cnt_out = (uint8_t)(c.plan_count % 100);
    c.plan_his_behavior = d_behavior;
    uint8_t cnt = (uint8_t)(c.plan_count + 1);
    if (cnt % 100 == 1) cnt = 1;
    c.plan_count = cnt;
    if (lane == 0) {
        dp_plan_record r;
        r.velocity_expect = v_exp; r.path_lat_dis = lat; r.path_dir_err = dir_err; r.remain_dis = remain;
        r.mindist_lat = ls.dis_lat; r.mindist_lon = ls.dis_lng; r.brakespeed = brake; r.des_acc = des_acc; r.radius = radius;
        r.aim_x = c.aim_x; r.aim_y = c.aim_y; r.aim_dir = c.aim_dir; r.aim_id = c.aim_id;
        r.behavior = c.behavior; r.target_roadnum = c.target_roadnum; r.target_lanenum = c.target_lanenum;
        r.light = c.light_status; r.behavior_to_dlg = c.behavior_to_dlg; r.afresh_cause = (uint16_t)cause;
        r.sweep_index = (int16_t)sweep_pick; r.path_near_id = (int16_t)near_id; r.path_front_near_id = (int16_t)front_id;
        r.ob_index = (int16_t)ls.ob; r.ob_pathid = (uint16_t)ls.pathid; r.n_traj = (uint16_t)n_traj;
        r.afresh_planning = afresh; r.ob_flag = ls.found; r.acc_flag = acc_flag; r.cnt = cnt_out;
        rec[scene] = r;
        carry[scene] = c;
        if (tr) { tr->ub_hits = (uint16_t)ub; tr->pts_scored = (uint32_t)pts; }
    }
}

// ---- reset carry to constructor state (Decision.cpp:8-29, Planning.cpp:8-11,62) ----
__global__ void dp_reset_kernel(dp_carry* carry, double* last_path, int first, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    dp_carry c;
    memset(&c, 0, sizeof(c));
    c.behavior = 1; c.velocity_expect = 10; c.his_behavior = 1; c.plan_his_behavior = 1;
    carry[first + i] = c;
    double* lp = last_path + (size_t)(first + i) * 2 * DP_PATH_POINTS;
    for (int k = 0; k < 2 * DP_PATH_POINTS; ++k) lp[k] = 0.0;
}

// ---- map precompute: per-point segment length (CalcDistance idiom) and unit right normal ----
__global__ void dp_map_prep_kernel(const double* x, const double* y, const int32_t* lane_pt_off, int n_lanes, double* nx, double* ny,
                                   double* lenp) {
    const int gl = blockIdx.x;
    if (gl >= n_lanes) return;
    const int off = lane_pt_off[gl], n = lane_pt_off[gl + 1] - off;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double vx = 0.0, vy = 0.0, l = 0.0;
        if (i + 1 < n) {
            const double2 nn = dp_normal(make_double2(x[off + i], y[off + i]), make_double2(x[off + i + 1], y[off + i + 1]));
            vx = nn.x; vy = nn.y;
            l = dp_dist_plain(x[off + i + 1], y[off + i + 1], x[off + i], y[off + i]);
        }
        nx[off + i] = vx; ny[off + i] = vy; lenp[off + i] = l;
    }
}

// ---- launchers (called from dp_api.cu) ----
cudaError_t dp_launch_cycle(const DevMap& m, const dp_params& p, int n, const dp_scene_hdr* hdr, const double* ox, const double* oy,
                            int max_obs, dp_carry* carry, double* last_path, dp_plan_record* rec, dp_trace_record* trace,
                            double* path_xy, double* path_ll, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int blocks = (n + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK;
    dp_cycle_kernel<<<blocks, DP_WARPS_PER_BLOCK * 32, 0, st>>>(m, p, n, hdr, ox, oy, max_obs, carry, last_path, rec, trace, path_xy, path_ll);
    return cudaGetLastError();
}
cudaError_t dp_launch_reset(dp_carry* carry, double* last_path, int first, int count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    dp_reset_kernel<<<(count + 127) / 128, 128, 0, st>>>(carry, last_path, first, count);
    return cudaGetLastError();
}
cudaError_t dp_launch_map_prep(const double* x, const double* y, const int32_t* lane_pt_off, int n_lanes, double* nx, double* ny,
                               double* lenp, cudaStream_t st) {
    dp_map_prep_kernel<<<n_lanes, 256, 0, st>>>(x, y, lane_pt_off, n_lanes, nx, ny, lenp);
    return cudaGetLastError();
}
