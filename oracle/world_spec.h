// oracle/world_spec.h -- TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
//
// Frozen arithmetic of (1) the world step of the closed-loop episode runner (include/dmpp_b200.h section 9) and (2) the
// output frames (section 8).  The reference has NO vehicle / traffic / localisation model -- those are other modules of its
// application (Decision.cpp:155-169 only reads their results) -- so the world step is this repository's own definition:
// "parity unpinned" for it means parity against this file only.  The frames restate Planning.cpp:173-214 and ARE pinned:
// oracle/ref_harness.cpp fills the same layout from the unmodified reference's PlanningOut / PlanningStatus objects.
//
// Arithmetic discipline as everywhere: IEEE binary64, operations in the order written, no fused multiply-add
// (-ffp-contract=off here, -fmad=false on the device), headings through spec::calc_global_dir.
#pragma once
#include "../include/dmpp_b200.h"
#include "planner_oracle.h"

namespace oracle {

void world_default_params(dp_world_params* p);

// one world step of one scene; rec == nullptr: place the agents and localise only.
// last_x / last_y: the carried local path (road_points of the cycle that produced rec).
void world_step(const MapView& m, const dp_params& p, const dp_world_params& wp, dp_scene_hdr& h, dp_agent* agents, double* ox,
                double* oy, const dp_plan_record* rec, const double* last_x, const double* last_y);

// Planning.cpp:173-214 in the frame layout of include/dmpp_b200.h; path = road_points of the cycle ([200] x, [200] y)
void pack_frames(const dp_params& p, const dp_plan_record& rec, const double* path_x, const double* path_y, dp_ctrl_frame* ctrl,
                 dp_status_frame* status);

}  // namespace oracle
