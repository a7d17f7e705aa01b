// host/PlannerBatch.h -- batched host mirror of the two reference classes: what `CDecision::Instance()` +
// `CPlanning::Instance()` (Decision.h:105-107, Planning.h:38-40) become when N independent scenes are planned at
// once.  One `Cycle()` = one iteration of CDecisionThread (Decision.cpp:119-206) + one of CPlanningThread
// (Planning.cpp:64-226) for every scene (one launch of the group kernel, or the overlapped Decision / Planning launch pair of
// the warp-per-scene kernel: csrc/dp_cycle.cu).  Thin, header-only wrapper of the C ABI; the class surface of the two
// reference singletons on top of it is in BatchFacade.h.  Compiled and run by tests/test_facade_cpp.py.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/dmpp_b200.h"

class CPlannerBatch {
public:
    CPlannerBatch(int max_scenes, int max_obs, int device = 0, const dp_params* params = nullptr)
        : max_scenes_(max_scenes), max_obs_(max_obs) {
        check(dp_create(&ctx_, device, params, max_scenes, max_obs), "dp_create");
    }
    ~CPlannerBatch() { dp_destroy(ctx_); }
    CPlannerBatch(const CPlannerBatch&) = delete;
    CPlannerBatch& operator=(const CPlannerBatch&) = delete;

    void UploadMap(const dp_map_desc& map) { check(dp_map_upload(ctx_, &map), "dp_map_upload"); }
    // constructor state of both singletons for scenes [first, first+count)  (Decision.cpp:8-29, Planning.cpp:8-11)
    void Reset(int first, int count) { check(dp_reset(ctx_, first, count), "dp_reset"); }
    // host buffers in, host buffers out; trace / paths may be null
    void Cycle(int first, int n, const dp_scene_hdr* hdr, const double* obs_x, const double* obs_y, dp_plan_record* rec,
               dp_trace_record* trace = nullptr, double* road_points = nullptr, double* latlng = nullptr) {
        check(dp_cycle_batch(ctx_, first, n, hdr, obs_x, obs_y, rec, trace, road_points, latlng), "dp_cycle_batch");
    }
    // device buffers, asynchronous on `stream`
    void CycleDevice(int first, int n, const dp_scene_hdr* hdr, const double* obs_x, const double* obs_y, dp_plan_record* rec,
                     void* stream, dp_trace_record* trace = nullptr, double* road_points = nullptr, double* latlng = nullptr) {
        check(dp_cycle_batch_dev(ctx_, first, n, hdr, obs_x, obs_y, rec, trace, road_points, latlng, stream), "dp_cycle_batch_dev");
    }
    // the two frames CPlanningThread publishes at the end of a cycle (Planning.cpp:173-214), for the cycle that just ran
    void PackFrames(int first, int n, const dp_plan_record* rec, dp_ctrl_frame* ctrl, dp_status_frame* status) {
        check(dp_pack_frames(ctx_, first, n, rec, ctrl, status), "dp_pack_frames");
    }
    // V2XEventDecision (Decision.cpp:283) for n scenes; the flags are the caller's to use, as in the reference
    void V2XEvents(int n, const dp_scene_hdr* hdr, const dp_v2x_data* v2x, const double* wp_lat, const double* wp_lng, int n_wp,
                   dp_v2x_flags* out, int mode = 0) {
        check(dp_v2x_event_batch(ctx_, n, hdr, v2x, wp_lat, wp_lng, n_wp, mode, out), "dp_v2x_event_batch");
    }
    dp_ctx* ctx() const { return ctx_; }
    int max_scenes() const { return max_scenes_; }
    int max_obs() const { return max_obs_; }

private:
    static void check(int rc, const char* what) {
        if (rc != DP_OK) throw std::runtime_error(std::string(what) + ": " + dp_last_error());
    }
    dp_ctx* ctx_ = nullptr;
    int max_scenes_, max_obs_;
};
