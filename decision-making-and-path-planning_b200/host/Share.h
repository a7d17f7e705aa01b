// host/Share.h -- drop-in replacement for the reference's `Share.h` (included by Planning.h:2 and
// Decision.h:2): the same type and macro names the two reference translation units use, and a `CShare`
// whose geometry operators run on the GPU through the C ABI of libdmpp_b200.so (include/dmpp_b200.h).
// With this header on the include path the UNMODIFIED Decision.cpp / Planning.cpp compile and every
// CShare::SearchObstacle / CreateNewPath / BezierPlanning / MeanPoints call they make is a CUDA launch
// (INTEGRATION.md).  The batched, fused entry points are in PlannerBatch.h.
#pragma once
#include "stdafx.h"                     // the application's own precompiled header (Win32 names, std headers)
#include "../../include/dmpp_b200.h"

#define REF_PATHPOINT DP_PATH_POINTS
#define LANESUM DP_LANESUM
#define ID_MORE 8
#define EPSILON 1e-6
#define PI 3.14159265358979323846
#define ROAD_FARAIM_MAX 60
#define ROAD_FARAIM_MIN 15
#define PRE_INTER_FARAIM 20
#define INTER_FARAIM 15
#define ROAD_REMAIN_DISTANCE 15
#define INTER_REMAIN_DISTANCE 5
#define Vehicle_Width 1.8
#define SYS_Frequency 1000.0

struct GlobalPoint2D { double x, y; };
struct GlobalPoint3D { double x, y, dir; };
struct GPSPoint2D { double lat, lng; };
struct GPSPoint3D { double lat, lng, ang; };
struct ObPoint { double x, y; int type; };
struct Obs_To_Veh { double dis_lat, dis_lng; };
struct Path_Obs { Obs_To_Veh Ob_Pose; bool Obs_flag; WORD Ob_Pathid; ObPoint Ob_Attr; };
struct Behavior_Dec { WORD behavior, target_lanenum, light_status; bool lanechg_status, obsavoid_status; WORD behavior_to_dlg; };
struct MapPoint { GlobalPoint3D global_point; WORD lane_sum, lane_width, lanechg_attribute; };
struct LocationOut {
    GlobalPoint3D globalpoint; GPSPoint2D gpspoint; int id[LANESUM];
    WORD lane_num, road_num, last_roadnum, next_roadnum, last_lanenum, next_lanenum, path_num;
    BYTE pos; double velocity; double period_last;
};
struct VehStatus { double reserved; };
struct PathInfo { WORD out_lane_no[LANESUM]; WORD stub_attribute; };
struct RoadInfo { int reserved; };
struct DecisionOut {
    double period_max, period_last; WORD behavior, target_roadnum, target_lanenum, light;
    double velocity_expect; vector<GlobalPoint2D> refpath; WORD behavior_to_dlg;
};
struct AimPoint { GlobalPoint3D Aim_point; INT Aim_id; };
struct PlanningOut {
    BYTE cnt, APA; double brakedis, brake_speed; bool desaccVd; double desacc, desspd, desstr; bool desstrVd;
    WORD light; double radius; BYTE road_type; BOOL sstop; GlobalPoint2D pnts[100];
};
struct PlanningStatus { int afresh_cause; double near_ob_dist, planspeed, planacc; WORD trafficlight; GlobalPoint2D path_points[100]; };
struct V2X_Data {
    double PedesDistance, PedesLatitude, PedesLongitude; int PedesDirection, SPATLaneOccupied, SPATState;
    double RSILatitude, RSILongitude; int V2XWarnStatus;
};
struct V2X_DataOut { int reserved; };
struct V2VWarn { int reserved; };
struct WarningPoint { double latitude, longitude; };

class CShare {
public:
    // GPU-backed (one launch per call): Planning.cpp:168,606,863,872; Decision.cpp:370,...,962
    bool SearchObstacle(vector<GlobalPoint2D> path, vector<ObPoint> obs, double lat_min, double lat_max,
                        double& dis_lat, double& dis_lng, ObPoint& ob, WORD& pathid);
    vector<GlobalPoint2D> CreateNewPath(vector<GlobalPoint2D> path, double offset);
    void BezierPlanning(GlobalPoint3D start, GlobalPoint3D aim, GlobalPoint2D out[], int n);
    void MeanPoints(GlobalPoint2D in[], int n_in, GlobalPoint2D out[], int n_out);
    // O(1) scalar helpers ("negligible"): evaluated inline on the host
    double CalcDistance(GlobalPoint2D a, GlobalPoint2D b);
    double CalcDistance(GPSPoint2D a, GPSPoint2D b);
    double CalcGlobalDir(GlobalPoint2D a, GlobalPoint2D b);
    int NearestId(GlobalPoint2D q, vector<GlobalPoint2D> pts);
    double LatDis(GlobalPoint2D q, GlobalPoint2D pt, GlobalPoint2D pt_next);
    GlobalPoint2D WGS84ToGlobal(GPSPoint2D g);
    GPSPoint2D GlobalToWGS84(GlobalPoint2D p);
    // the process-wide device context the single-call operators use (created on first use, device 0)
    static dp_ctx* Context();
    static long long SearchCalls();
};
