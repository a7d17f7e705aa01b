// oracle/planner_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
//
// Re-entrant, explicit-state CPU restatement of one Decision cycle (Decision.cpp:119-206 with
// SegmentDecision :216-315, PreStubDecision :323-402, StubDecision :409-486) followed by one
// Planning cycle (Planning.cpp:64-226) of the reference.  All cross-cycle state the reference
// hides in function statics and singleton members is carried in
// SceneState, so scenes are independent and can run on any number of host threads.
// The restatement is differential-tested against the UNMODIFIED reference objects
// (oracle/_ref/libref.so, tests/test_oracle_vs_ref.py); the geometry operators it calls are the
// frozen specification in cshare_spec.h ("parity unpinned" -- see that header).
#pragma once
#include <vector>
#include "../include/dmpp_b200.h"
#include "cshare_spec.h"
#include "ref_api.h"

#include <atomic>

namespace oracle {

// hit counters of the right-lane-change sites (Decision.cpp:1108,1382,1618,1711; Planning.cpp:473-501)
enum { BR_B3_NAV_1108, BR_B3_OBS_1382, BR_B3_BOTH_1618, BR_B3_BOTH_1711, BR_ENTER_1596, BR_ENTER_1688, BR_AIM_RIGHT_473,
       BR_AIM_RIGHT_WALK, BR_COUNT };
extern std::atomic<long long> g_branch_hits[BR_COUNT];

struct MapView {
    dp_map_desc d;
    int lanes_of(int road) const { return d.road_lane_base[road] - d.road_lane_base[road - 1]; }
    int lane_index(int road, int lane) const { return d.road_lane_base[road - 1] + lane - 1; }
    int lane_size(int gl) const { return d.lane_pt_off[gl + 1] - d.lane_pt_off[gl]; }
    spec::P2 pt(int gl, int i) const { int k = d.lane_pt_off[gl] + i; return spec::P2{d.x[k], d.y[k]}; }
    double dir(int gl, int i) const { return d.dir[d.lane_pt_off[gl] + i]; }
    int width(int gl, int i) const { return d.lane_width[d.lane_pt_off[gl] + i]; }
    int attr(int gl, int i) const { return d.lanechg_attr[d.lane_pt_off[gl] + i]; }
};

struct SceneState {
    dp_carry c;
    double last_x[DP_PATH_POINTS], last_y[DP_PATH_POINTS];   // CPlanning::last_Bpoints (Planning.cpp:6)
};

void reset_state(SceneState& s);   // constructor state: Decision.cpp:8-29, Planning.cpp:8-11,62

struct CycleOut {
    dp_plan_record* rec;           // required
    dp_trace_record* trace;        // nullable
    double* path_xy;               // nullable [2][200]
    double* path_ll;               // nullable [2][100]
    ref_call* calls;               // nullable, call log in reference call order
    int calls_cap;
    int n_calls;                   // out
    int ub_hits;                   // out: reference UB sites the restatement had to define (should stay 0)
    // in (BASELINE config 5 generalisation, nullable): the scene's [T x n_obs] predicted track tile; the junction search
    // (Decision.cpp:370, :455) then runs against moving agents (spec::search_obstacle_tile).  Everything else stays static.
    const double* tile_x = nullptr; const double* tile_y = nullptr; int tile_T = 0;
};

// exhaustive_sweep: also evaluate the avoid candidates the reference skips after its `break`
// (fills trace->sweep completely; does not change any decision).
void cycle(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const double* ox, const double* oy,
           SceneState& st, CycleOut& out, bool exhaustive_sweep);

}  // namespace oracle
