// oracle/compat/GAC_Auotpilot_DP.h -- TEST INFRASTRUCTURE. Stand-in for the MFC application
// object that owns all shared state in the reference (Planning.cpp:3,47; Decision.cpp:3,80):
// 4 auto-reset events, 5 critical sections, the HD-map tables and the Get*/Set* accessors the two
// threads call.  oracle/ref_harness.cpp fills it per cycle and captures what the threads publish.
#pragma once
#include <Share.h>   // angle form: the -I order picks oracle/compat/Share.h (CPU spec) or host/Share.h (GPU)

typedef vector<MapPoint> LanePts;                       // [id]
typedef vector<vector<LanePts>> RoadMap;                // [road-1][lane-1][id]
typedef vector<vector<vector<vector<LanePts>>>> InterMap;  // [last_road-1][next_road-1][last_lane-1][next_lane-1][id]

class CGAC_Auotpilot_DPApp {
public:
    HANDLE x_PercetionPreProcessingEvent, x_LocationEvent, x_DecisionEvent, x_PlanningEvent;
    CCriticalSection x_criticalDecision, x_criticalLocation, x_criticalVhclHisPos, x_criticalObstacle,
        x_criticalV2X_Data;
    RoadMap decision_MapData, planning_MapData;
    InterMap decision_InterMapData, planning_InterMapData;

    // inputs published by the harness
    LocationOut in_location;
    VehStatus in_vehstatus;
    vector<ObPoint> in_obs;
    vector<RoadInfo> in_roadinfo;
    vector<PathInfo> in_roadnavi;
    V2X_Data in_v2x;
    vector<WarningPoint> in_rsi;
    // outputs captured from the threads
    DecisionOut out_decision;
    PlanningStatus out_status;
    PlanningOut out_planning;
    int n_set_decision, n_set_planning;

    DecisionOut GetDesicionOut() { return out_decision; }
    LocationOut GetLocationOut() { return in_location; }
    VehStatus GetVehStatus() { return in_vehstatus; }
    vector<ObPoint> GetObj() { return in_obs; }
    vector<RoadInfo> GetRoadInfo() { return in_roadinfo; }
    vector<PathInfo> GetRoadNavi() { return in_roadnavi; }
    V2X_Data GetV2XData() { return in_v2x; }
    vector<WarningPoint> GetRSIWarningPointList() { return in_rsi; }
    void SetDecisionOut(const DecisionOut& d) { out_decision = d; ++n_set_decision; }
    void SetPlanningStatus(const PlanningStatus& s) { out_status = s; }
    void SetUdpSendCtrl(const PlanningOut& p) { out_planning = p; ++n_set_planning; }
};
