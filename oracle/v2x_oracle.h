// oracle/v2x_oracle.h -- TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
// Restatement of the V2X event handlers of the reference (Decision.cpp:1824-2434) on explicit inputs; pinned to the unmodified
// reference's own private methods by tests/test_v2x.py (oracle/ref_harness.cpp: ref_v2x_event).
#pragma once
#include "../include/dmpp_b200.h"
#include "planner_oracle.h"

namespace oracle {
// one scene; mode as in dp_v2x_event_batch
void v2x_event(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const dp_v2x_data& v, const double* wp_lat,
               const double* wp_lng, int mode, dp_v2x_flags* out);
// the opt-in speed command of include/dmpp_b200.h section 10 (an extension: the reference drops the flags)
void v2x_apply(const dp_v2x_flags& f, dp_plan_record& r);
}  // namespace oracle
