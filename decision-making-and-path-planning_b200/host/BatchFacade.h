// host/BatchFacade.h -- the PUBLIC surface of the reference's two classes, kept, on top of the batched C ABI.
//
// Reference surface (Decision.h:105-107, Planning.h:38-85):
//     static CDecision& Instance();   BYTE startCDecisionThread();
//     static CPlanning& Instance();   BYTE startCPlanningThread();
//     public CPlanning fields: path_lat_dis, afresh_planning, afresh_cause, remain_dis, path_dir_err, path_near_id,
//                              path_front_near_id, his_behavior, brakespeed, acc_flag, des_acc
// In the reference each singleton owns one thread that runs one scene; here the same two singletons own one batch of scenes
// and one `Cycle()` of the shared batch is "one wake-up of CDecisionThread followed by one of CPlanningThread" for all of them
// (Decision.cpp:119-206, Planning.cpp:64-226).  The Win32 thread + event shells are the host
// application's business: start*Thread() only binds the singleton to the batch and returns 1 like the reference
// (Decision.cpp:45-53), the application calls Cycle() from whatever loop it has.  Every per-scene value the reference
// publishes is read back through accessors named after the reference members.
#pragma once
#include "PlannerBatch.h"

namespace dmpp {

typedef unsigned char BYTE;
typedef unsigned short WORD;

// the batch both singletons share (the reference's two threads share the application object, Decision.cpp:80, Planning.cpp:47)
class CBatchApp {
public:
    static CBatchApp& Instance() { static CBatchApp a; return a; }
    // replaces the application start-up that loads the map and creates the two objects
    void Open(int max_scenes, int max_obs, const dp_map_desc& map, int device = 0, const dp_params* params = nullptr) {
        batch_.reset(new CPlannerBatch(max_scenes, max_obs, device, params));
        batch_->UploadMap(map);
        rec_.assign((size_t)max_scenes, dp_plan_record());
        n_ = 0;
    }
    void Close() { batch_.reset(); }
    bool IsOpen() const { return (bool)batch_; }
    // one Decision + Planning cycle of scenes [0, n): LocationOut / RoadNavi slices in hdr, obstacle points in obs_x / obs_y
    void Cycle(int n, const dp_scene_hdr* hdr, const double* obs_x, const double* obs_y) {
        if (!batch_) throw std::runtime_error("CBatchApp::Cycle: Open() first");
        if (!decision_started_ || !planning_started_) throw std::runtime_error("CBatchApp::Cycle: start both threads first");
        batch_->Cycle(0, n, hdr, obs_x, obs_y, rec_.data());
        n_ = n;
        if (publish_frames_) {                               // app->SetPlanningStatus / app->SetUdpSendCtrl (Planning.cpp:186, :214)
            ctrl_.resize((size_t)n); status_.resize((size_t)n);
            batch_->PackFrames(0, n, rec_.data(), ctrl_.data(), status_.data());
        }
    }
    // what the reference hands to the controller link and to the debug UI at the end of every Planning cycle
    void PublishFrames(bool on) { publish_frames_ = on; }
    const dp_ctrl_frame& UdpSendCtrl(int scene) const {
        if (!publish_frames_ || scene < 0 || scene >= n_) throw std::out_of_range("frame");
        return ctrl_[(size_t)scene];
    }
    const dp_status_frame& PlanningStatus(int scene) const {
        if (!publish_frames_ || scene < 0 || scene >= n_) throw std::out_of_range("frame");
        return status_[(size_t)scene];
    }
    const dp_plan_record& Record(int scene) const {
        if (scene < 0 || scene >= n_) throw std::out_of_range("scene index");
        return rec_[(size_t)scene];
    }
    int scenes() const { return n_; }
    bool decision_started_ = false, planning_started_ = false;

private:
    CBatchApp() {}
    std::unique_ptr<CPlannerBatch> batch_;
    std::vector<dp_plan_record> rec_;
    std::vector<dp_ctrl_frame> ctrl_;
    std::vector<dp_status_frame> status_;
    bool publish_frames_ = false;
    int n_ = 0;
};

class CDecision {
public:
    static CDecision& Instance() { static CDecision d; return d; }                     // Decision.cpp:36-40
    BYTE startCDecisionThread() {                                                       // Decision.cpp:45-53: 1 = started
        if (!CBatchApp::Instance().IsOpen()) return 0;
        CBatchApp::Instance().decision_started_ = true;
        return 1;
    }
    // DecisionOut of one scene after the last cycle (Decision.cpp:187-196)
    WORD behavior(int s) const { return CBatchApp::Instance().Record(s).behavior; }
    WORD target_roadnum(int s) const { return CBatchApp::Instance().Record(s).target_roadnum; }
    WORD target_lanenum(int s) const { return CBatchApp::Instance().Record(s).target_lanenum; }
    WORD light(int s) const { return CBatchApp::Instance().Record(s).light; }
    double velocity_expect(int s) const { return CBatchApp::Instance().Record(s).velocity_expect; }
    WORD behavior_to_dlg(int s) const { return CBatchApp::Instance().Record(s).behavior_to_dlg; }

private:
    CDecision() {}
};

class CPlanning {
public:
    static CPlanning& Instance() { static CPlanning p; return p; }                     // Planning.cpp:18-22
    BYTE startCPlanningThread() {                                                       // Planning.cpp:27-35
        if (!CBatchApp::Instance().IsOpen()) return 0;
        CBatchApp::Instance().planning_started_ = true;
        return 1;
    }
    // the public fields of Planning.h:41-51, per scene
    double path_lat_dis(int s) const { return CBatchApp::Instance().Record(s).path_lat_dis; }
    bool afresh_planning(int s) const { return CBatchApp::Instance().Record(s).afresh_planning != 0; }
    int afresh_cause(int s) const { return CBatchApp::Instance().Record(s).afresh_cause; }
    double remain_dis(int s) const { return CBatchApp::Instance().Record(s).remain_dis; }
    double path_dir_err(int s) const { return CBatchApp::Instance().Record(s).path_dir_err; }
    int path_near_id(int s) const { return CBatchApp::Instance().Record(s).path_near_id; }
    int path_front_near_id(int s) const { return CBatchApp::Instance().Record(s).path_front_near_id; }
    double brakespeed(int s) const { return CBatchApp::Instance().Record(s).brakespeed; }
    bool acc_flag(int s) const { return CBatchApp::Instance().Record(s).acc_flag != 0; }
    double des_acc(int s) const { return CBatchApp::Instance().Record(s).des_acc; }
    // PlanningOut scalars (Planning.cpp:189-201)
    double brakedis(int s) const { return CBatchApp::Instance().Record(s).mindist_lon; }
    double radius(int s) const { return CBatchApp::Instance().Record(s).radius; }

private:
    CPlanning() {}
};

}  // namespace dmpp
