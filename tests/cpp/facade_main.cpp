// tests/cpp/facade_main.cpp -- TEST: the batched class facade of the product (host/BatchFacade.h: dmpp::CDecision /
// dmpp::CPlanning over CPlannerBatch over the C ABI) must publish, scene by scene and cycle by cycle, what the oracle publishes.
// Input: a dump written by tests/test_facade_cpp.py (map tables, scene headers, obstacle rows, the oracle's plan records).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../decision-making-and-path-planning_b200/host/BatchFacade.h"

template <class T> static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}
static bool same(double a, double b) { return (a == b) || (a != a && b != b); }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: facade_main dump.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("dump"); return 2; }
    int32_t hd[8];
    if (fread(hd, 4, 8, f) != 8) return 2;
    const int n_roads = hd[0], n_lanes = hd[1], n_conn = hd[2], n_points = hd[3], n = hd[4], cycles = hd[5], max_obs = hd[6];
    auto rlb = rd<int32_t>(f, n_roads + 1); auto lpo = rd<int32_t>(f, n_lanes + 1); auto conn = rd<dp_connector>(f, n_conn);
    auto x = rd<double>(f, n_points); auto y = rd<double>(f, n_points); auto dir = rd<double>(f, n_points);
    auto width = rd<uint16_t>(f, n_points); auto attr = rd<uint16_t>(f, n_points);
    auto H = rd<dp_scene_hdr>(f, (size_t)cycles * n); auto OX = rd<double>(f, (size_t)cycles * n * max_obs);
    auto OY = rd<double>(f, (size_t)cycles * n * max_obs); auto want = rd<dp_plan_record>(f, (size_t)cycles * n);
    auto want_ctrl = rd<dp_ctrl_frame>(f, (size_t)cycles * n); auto want_status = rd<dp_status_frame>(f, (size_t)cycles * n);
    // V2X events for the scenes of cycle 0 and the oracle's flags in both modes
    const int n_wp = hd[7];
    auto v2x = rd<dp_v2x_data>(f, (size_t)n); auto wp_lat = rd<double>(f, (size_t)n_wp); auto wp_lng = rd<double>(f, (size_t)n_wp);
    auto want_f0 = rd<dp_v2x_flags>(f, (size_t)n); auto want_f1 = rd<dp_v2x_flags>(f, (size_t)n);
    fclose(f);
    dp_map_desc md;
    md.n_roads = n_roads; md.road_lane_base = rlb.data(); md.n_lanes = n_lanes; md.lane_pt_off = lpo.data(); md.n_conn = n_conn;
    md.conn = conn.data(); md.n_points = n_points; md.x = x.data(); md.y = y.data(); md.dir = dir.data(); md.lane_width = width.data();
    md.lanechg_attr = attr.data();

    using namespace dmpp;
    if (CDecision::Instance().startCDecisionThread() != 0) { fprintf(stderr, "thread started without an application\n"); return 1; }
    CBatchApp::Instance().Open(n, max_obs, md);
    if (CDecision::Instance().startCDecisionThread() != 1 || CPlanning::Instance().startCPlanningThread() != 1) return 1;
    CBatchApp::Instance().PublishFrames(true);
    CDecision& D = CDecision::Instance();
    CPlanning& P = CPlanning::Instance();
    long bad = 0, b3 = 0;
    for (int c = 0; c < cycles; ++c) {
        CBatchApp::Instance().Cycle(n, &H[(size_t)c * n], &OX[(size_t)c * n * max_obs], &OY[(size_t)c * n * max_obs]);
        for (int s = 0; s < n; ++s) {
            const dp_plan_record& w = want[(size_t)c * n + s];
            bool ok = D.behavior(s) == w.behavior && D.target_lanenum(s) == w.target_lanenum && D.light(s) == w.light &&
                      D.behavior_to_dlg(s) == w.behavior_to_dlg && same(D.velocity_expect(s), w.velocity_expect) &&
                      same(P.path_lat_dis(s), w.path_lat_dis) && P.afresh_planning(s) == (w.afresh_planning != 0) &&
                      P.afresh_cause(s) == w.afresh_cause && same(P.remain_dis(s), w.remain_dis) && P.path_near_id(s) == w.path_near_id &&
                      P.path_front_near_id(s) == w.path_front_near_id && same(P.brakespeed(s), w.brakespeed) &&
                      P.acc_flag(s) == (w.acc_flag != 0) && same(P.des_acc(s), w.des_acc) && same(P.brakedis(s), w.mindist_lon) &&
                      same(P.radius(s), w.radius) && !(P.path_dir_err(s) - w.path_dir_err > 1e-9) && !(w.path_dir_err - P.path_dir_err(s) > 1e-9);
            ok = ok && memcmp(&CBatchApp::Instance().UdpSendCtrl(s), &want_ctrl[(size_t)c * n + s], sizeof(dp_ctrl_frame)) == 0 &&
                 memcmp(&CBatchApp::Instance().PlanningStatus(s), &want_status[(size_t)c * n + s], sizeof(dp_status_frame)) == 0;
            if (!ok && bad++ < 5) fprintf(stderr, "mismatch cycle %d scene %d: behavior %d vs %d\n", c, s, (int)D.behavior(s), (int)w.behavior);
            b3 += D.behavior(s) != 1;
        }
    }
    {   // the V2X handlers through the same C++ wrapper (its own context: the facade keeps its batch private)
        CPlannerBatch vb(n, max_obs);
        vb.UploadMap(md);
        std::vector<dp_v2x_flags> got((size_t)n);
        for (int mode = 0; mode < 2; ++mode) {
            vb.V2XEvents(n, H.data(), v2x.data(), wp_lat.data(), wp_lng.data(), n_wp, got.data(), mode);
            const std::vector<dp_v2x_flags>& want_f = mode ? want_f1 : want_f0;
            for (int s = 0; s < n; ++s)
                if (memcmp(&got[(size_t)s], &want_f[(size_t)s], sizeof(dp_v2x_flags)) != 0 && bad++ < 5)
                    fprintf(stderr, "V2X mismatch mode %d scene %d\n", mode, s);
        }
    }
    CBatchApp::Instance().Close();
    printf("facade: %d scenes x %d cycles, %ld mismatches, %ld non-keep decisions\n", n, cycles, bad, b3);
    if (bad == 0) printf("FACADE OK\n");
    return bad == 0 ? 0 : 1;
}
