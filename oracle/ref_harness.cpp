// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE. Drives the UNMODIFIED reference translation
// units (/root/reference/Decision.cpp, Planning.cpp, compiled where they lie by oracle/Makefile
// into oracle/_ref/libref.so) through whole episodes:
//   * CreateThread / WaitForSingleObject / SetEvent are implemented as cooperative coroutines
//     (ucontext), so the reference's own thread loops (Decision.cpp:119-206, Planning.cpp:64-226)
//     run verbatim, one iteration per published cycle, Decision then Planning;
//   * QueryPerformanceCounter is a deterministic clock advanced by the scripted cycle period;
//   * the app object (compat/GAC_Auotpilot_DP.h) is filled from the same dp_scene_hdr / obstacle
//     arrays the CUDA path consumes, and everything the threads publish is captured.
// Function-local statics of BehaviorDecision (Decision.cpp:915-917) are reset between episodes
// through symbols globalised by objcopy in the Makefile (the sources are not edited).
#include <time.h>
#include <ucontext.h>
#include "Decision.h"   // from /root/reference via -I
#include "Planning.h"
#include "GAC_Auotpilot_DP.h"
#include <new>
#include "../include/dmpp_b200.h"
#include "ref_api.h"

static CGAC_Auotpilot_DPApp g_app;
void* AfxGetApp() { return &g_app; }
static int g_msgbox = 0;
int AfxMessageBox(const char*) { ++g_msgbox; return 0; }

// ---- function statics of CDecision::BehaviorDecision, globalised by objcopy ------------------
extern UINT ref_frontobs_time asm(
    "_ZZN9CDecision16BehaviorDecisionE11LocationOutSt6vectorI8PathInfoSaIS2_EEjjdS1_I13GlobalPoint2DSaIS5_EERK8Path_ObsSA_SA_SA_SA_SA_12Behavior_DecRSB_E13frontobs_time");
extern UINT ref_obsavoid_time asm(
    "_ZZN9CDecision16BehaviorDecisionE11LocationOutSt6vectorI8PathInfoSaIS2_EEjjdS1_I13GlobalPoint2DSaIS5_EERK8Path_ObsSA_SA_SA_SA_SA_12Behavior_DecRSB_E13obsavoid_time");
extern UINT ref_no_obsaviod_time asm(
    "_ZZN9CDecision16BehaviorDecisionE11LocationOutSt6vectorI8PathInfoSaIS2_EEjjdS1_I13GlobalPoint2DSaIS5_EERK8Path_ObsSA_SA_SA_SA_SA_12Behavior_DecRSB_E16no_obsaviod_time");

// ---- cooperative threads ----------------------------------------------------------------------
namespace {
struct Co {
    ucontext_t ctx;
    char* stack = nullptr;
    LPTHREAD_START_ROUTINE fn = nullptr;
    LPVOID arg = nullptr;
    bool live = false;
};
Co g_co[2];
int g_ncos = 0;
int g_running = -1;       // index of the coroutine currently executing, -1 = harness
ucontext_t g_main;
int g_ev[4];              // perception, location, decision, planning
long long g_clock = 0;    // ticks; SYS_Frequency = 1000 ticks per second
const size_t kStack = 1 << 20;

void co_entry(int idx) {
    g_co[idx].fn(g_co[idx].arg);
    g_co[idx].live = false;
    swapcontext(&g_co[idx].ctx, &g_main);
}
void co_resume(int idx) {
    g_running = idx;
    swapcontext(&g_main, &g_co[idx].ctx);
    g_running = -1;
}
void co_drop_all() {
    for (int i = 0; i < 2; ++i) {
        free(g_co[i].stack);
        g_co[i] = Co();
    }
    g_ncos = 0;
}
}  // namespace

extern "C" HANDLE CreateThread(void*, size_t, LPTHREAD_START_ROUTINE fn, LPVOID arg, DWORD, DWORD* id) {
    if (g_ncos >= 2) return nullptr;
    Co& c = g_co[g_ncos];
    c.stack = (char*)malloc(kStack);
    c.fn = fn;
    c.arg = arg;
    c.live = true;
    getcontext(&c.ctx);
    c.ctx.uc_stack.ss_sp = c.stack;
    c.ctx.uc_stack.ss_size = kStack;
    c.ctx.uc_link = &g_main;
    makecontext(&c.ctx, (void (*)())co_entry, 1, g_ncos);
    if (id) *id = (DWORD)g_ncos;
    ++g_ncos;
    return (HANDLE)&c;
}
extern "C" DWORD WaitForSingleObject(HANDLE h, DWORD) {
    int* f = (int*)h;
    while (!*f) swapcontext(&g_co[g_running].ctx, &g_main);   // yield until the harness sets it
    *f = 0;                                                    // auto-reset
    return 0;
}
extern "C" BOOL SetEvent(HANDLE h) { *(int*)h = 1; return 1; }
extern "C" BOOL QueryPerformanceCounter(LARGE_INTEGER* t) { t->QuadPart = g_clock; return 1; }
extern "C" void Sleep(DWORD) {}

// ---- call log filled by share_impl.cpp (CPU spec build); the GPU-Share build counts calls only ----
#ifdef REF_GPU_SHARE
static ref_call* g_calllog = nullptr;
static int g_calllog_n = 0, g_calllog_cap = 0;
#define g_search_calls (CShare::SearchCalls())
extern "C" void share_set_datum(double, double, double, double) {}
#else
extern "C" ref_call* g_calllog;
extern "C" int g_calllog_n, g_calllog_cap;
extern "C" long long g_search_calls;
#endif

// ---- map ------------------------------------------------------------------------------------------
static dp_map_desc g_desc;
extern "C" const dp_map_desc* ref_map_desc() { return &g_desc; }   // (the caller keeps the arrays alive)
// capture of the frames the Planning thread publishes (ref_capture_frames), and a per-cycle hook (closed loop, ref_closed_loop.cpp)
static dp_ctrl_frame* g_cap_ctrl = nullptr;
static dp_status_frame* g_cap_status = nullptr;
static dp_ctrl_frame* g_cap_ctrl_base = nullptr;
static dp_status_frame* g_cap_status_base = nullptr;
static long g_cap_stride = 0;
static void (*g_hook)(int, void*) = nullptr;
static void* g_hook_user = nullptr;
extern "C" void ref_capture_frames(dp_ctrl_frame* ctrl, dp_status_frame* status) { g_cap_ctrl_base = ctrl; g_cap_status_base = status; }
extern "C" void ref_set_cycle_hook(void (*fn)(int, void*), void* user) { g_hook = fn; g_hook_user = user; }
extern "C" void ref_last_path(double* x, double* y) {
    for (int i = 0; i < 200; ++i) { x[i] = CPlanning::last_Bpoints[i].x; y[i] = CPlanning::last_Bpoints[i].y; }
}
// the reference's own output objects, field by field, into the frame layout of include/dmpp_b200.h
static void fill_frames(const PlanningOut& po, const PlanningStatus& ps, dp_ctrl_frame* cf, dp_status_frame* sf) {
    if (cf) {
        memset(cf, 0, sizeof(*cf));
        cf->cnt = po.cnt; cf->apa = po.APA; cf->desacc_vd = po.desaccVd; cf->desstr_vd = po.desstrVd; cf->road_type = po.road_type;
        cf->sstop = (uint8_t)po.sstop; cf->light = po.light; cf->brakedis = po.brakedis; cf->brake_speed = po.brake_speed;
        cf->desacc = po.desacc; cf->desspd = po.desspd; cf->desstr = po.desstr; cf->radius = po.radius;
        for (int i = 0; i < 100; ++i) { cf->pnts[i][0] = po.pnts[i].x; cf->pnts[i][1] = po.pnts[i].y; }
    }
    if (sf) {
        memset(sf, 0, sizeof(*sf));
        sf->afresh_cause = ps.afresh_cause; sf->trafficlight = ps.trafficlight; sf->near_ob_dist = ps.near_ob_dist;
        sf->planspeed = ps.planspeed; sf->planacc = ps.planacc;
        for (int i = 0; i < 100; ++i) { sf->path_points[i][0] = ps.path_points[i].x; sf->path_points[i][1] = ps.path_points[i].y; }
    }
}

extern "C" int ref_set_map(const dp_map_desc* m) {
    g_desc = *m;
    g_app.decision_MapData.clear();
    g_app.decision_MapData.resize(m->n_roads);
    for (int r = 0; r < m->n_roads; ++r) {
        int l0 = m->road_lane_base[r], l1 = m->road_lane_base[r + 1];
        g_app.decision_MapData[r].resize(l1 - l0);
        for (int l = l0; l < l1; ++l) {
            LanePts& v = g_app.decision_MapData[r][l - l0];
            for (int i = m->lane_pt_off[l]; i < m->lane_pt_off[l + 1]; ++i) {
                MapPoint p;
                p.global_point.x = m->x[i];
                p.global_point.y = m->y[i];
                p.global_point.dir = m->dir[i];
                p.lane_sum = (WORD)(l1 - l0);
                p.lane_width = m->lane_width[i];
                p.lanechg_attribute = m->lanechg_attr[i];
                v.push_back(p);
            }
        }
    }
    g_app.planning_MapData = g_app.decision_MapData;
    int R = m->n_roads, L = DP_LANESUM;
    InterMap im(R, vector<vector<vector<LanePts>>>(R, vector<vector<LanePts>>(L, vector<LanePts>(L))));
    for (int c = 0; c < m->n_conn; ++c) {
        const dp_connector& k = m->conn[c];
        LanePts& v = im[k.last_road - 1][k.next_road - 1][k.last_lane - 1][k.next_lane - 1];
        for (int i = m->lane_pt_off[k.lane]; i < m->lane_pt_off[k.lane + 1]; ++i) {
            MapPoint p;
            p.global_point.x = m->x[i];
            p.global_point.y = m->y[i];
            p.global_point.dir = m->dir[i];
            p.lane_sum = 1;
            p.lane_width = m->lane_width[i];
            p.lanechg_attribute = m->lanechg_attr[i];
            v.push_back(p);
        }
    }
    g_app.decision_InterMapData = im;
    g_app.planning_InterMapData = im;
    return 0;
}

// ---- one episode -----------------------------------------------------------------------------------
static void reset_singletons() {
    CDecision& d = CDecision::Instance();
    d.~CDecision();
    memset((void*)&d, 0, sizeof(d));
    new (&d) CDecision();
    CPlanning& p = CPlanning::Instance();
    p.~CPlanning();
    memset((void*)&p, 0, sizeof(p));
    new (&p) CPlanning();
    ref_frontobs_time = 0;
    ref_obsavoid_time = 0;
    ref_no_obsaviod_time = 0;
}

// hdr/obs are strided so that the same [cycle][scene] arrays that feed the GPU can be walked:
// element (cycle c) of this scene is hdr[c*hdr_stride], obs_x[c*obs_stride + k].
extern "C" int ref_run_episode(int cycles, const dp_scene_hdr* hdr, long hdr_stride, const double* obs_x,
                               const double* obs_y, long obs_stride, dp_plan_record* rec, long rec_stride,
                               double* path_xy, long path_xy_stride, double* path_ll, long path_ll_stride,
                               ref_call* calls, int32_t* n_calls, long ncalls_stride, long calls_stride, int calls_cap,
                               dp_carry* carry_out, double* last_path_out) {
    co_drop_all();
    reset_singletons();
    g_msgbox = 0;
    memset(g_ev, 0, sizeof(g_ev));
    g_app.x_PercetionPreProcessingEvent = &g_ev[0];
    g_app.x_LocationEvent = &g_ev[1];
    g_app.x_DecisionEvent = &g_ev[2];
    g_app.x_PlanningEvent = &g_ev[3];
    g_app.n_set_decision = g_app.n_set_planning = 0;
    memset(&g_app.in_v2x, 0, sizeof(g_app.in_v2x));
    g_app.in_rsi.clear();
    g_app.out_decision = DecisionOut();
    g_clock = 0;

    // RoadNavi is read once before the decision loop (Decision.cpp:85): assemble it from the
    // per-cycle slices of the whole episode.
    int max_path = 0;
    for (int c = 0; c < cycles; ++c)
        if (hdr[c * hdr_stride].path_num > max_path) max_path = hdr[c * hdr_stride].path_num;
    g_app.in_roadnavi.assign(max_path + 1, PathInfo());
    for (int c = 0; c < cycles; ++c) {
        const dp_scene_hdr& h = hdr[c * hdr_stride];
        PathInfo pi;
        for (int i = 0; i < LANESUM; ++i) pi.out_lane_no[i] = h.out_lane_no[i];
        pi.stub_attribute = h.stub_attribute;
        g_app.in_roadnavi[h.path_num] = pi;
    }

    if (!CDecision::Instance().startCDecisionThread()) return -1;
    if (!CPlanning::Instance().startCPlanningThread()) return -1;
    co_resume(0);   // runs the prologue of CDecisionThread up to its first wait
    co_resume(1);   // runs the prologue of CPlanningThread up to its first wait

    for (int c = 0; c < cycles; ++c) {
        const dp_scene_hdr& h = hdr[c * hdr_stride];
        long long ticks = (long long)llround(h.period_ms);
        if ((double)ticks / SYS_Frequency * 1000 != h.period_ms) return -2;   // period must be exactly reproducible
        g_clock += ticks;
        LocationOut& lo = g_app.in_location;
        memset(&lo, 0, sizeof(lo));
        lo.globalpoint.x = h.x; lo.globalpoint.y = h.y; lo.globalpoint.dir = h.dir;
        lo.velocity = h.velocity;
        for (int i = 0; i < LANESUM; ++i) lo.id[i] = h.id[i];
        lo.lane_num = h.lane_num; lo.road_num = h.road_num; lo.pos = (BYTE)h.pos; lo.path_num = h.path_num;
        lo.last_roadnum = h.last_roadnum; lo.next_roadnum = h.next_roadnum;
        lo.last_lanenum = h.last_lanenum; lo.next_lanenum = h.next_lanenum;
        g_app.in_obs.resize(h.n_obs);
        for (int k = 0; k < h.n_obs; ++k) {
            g_app.in_obs[k].x = obs_x[c * obs_stride + k];
            g_app.in_obs[k].y = obs_y[c * obs_stride + k];
            g_app.in_obs[k].type = 1;
        }
        g_calllog = calls ? calls + c * calls_stride : nullptr;
        g_calllog_n = 0;
        g_calllog_cap = calls_cap;
        int nd = g_app.n_set_decision, np = g_app.n_set_planning;
        const long long calls_before = g_search_calls; (void)calls_before;
        g_ev[0] = 1; g_ev[1] = 1;
        co_resume(0);                         // one Decision cycle; ends with SetEvent(x_DecisionEvent)
        if (g_app.n_set_decision != nd + 1) return -3;
        co_resume(1);                         // one Planning cycle
        if (g_app.n_set_planning != np + 1) return -4;
        g_ev[3] = 0;
#ifdef REF_GPU_SHARE
        g_calllog_n = (int)(g_search_calls - calls_before);
#endif
        if (n_calls) n_calls[c * ncalls_stride] = g_calllog_n;

        CPlanning& p = CPlanning::Instance();
        const DecisionOut& d = g_app.out_decision;
        const PlanningOut& po = g_app.out_planning;
        dp_plan_record& r = rec[c * rec_stride];
        memset(&r, 0, sizeof(r));
        r.velocity_expect = d.velocity_expect;
        r.path_lat_dis = p.path_lat_dis; r.path_dir_err = p.path_dir_err; r.remain_dis = p.remain_dis;
        r.mindist_lon = po.brakedis;
        r.mindist_lat = 0;                    // not published by the thread (local at Planning.cpp:161)
        r.brakespeed = p.brakespeed; r.des_acc = p.des_acc; r.radius = po.radius;
        r.aim_x = p.aimpoint_far.Aim_point.x; r.aim_y = p.aimpoint_far.Aim_point.y; r.aim_dir = p.aimpoint_far.Aim_point.dir;
        r.aim_id = p.aimpoint_far.Aim_id;
        r.behavior = d.behavior; r.target_roadnum = d.target_roadnum; r.target_lanenum = d.target_lanenum;
        r.light = d.light; r.behavior_to_dlg = d.behavior_to_dlg;
        r.afresh_cause = (uint16_t)p.afresh_cause;
        r.sweep_index = -2;                   // not observable in the reference
        r.path_near_id = (int16_t)p.path_near_id; r.path_front_near_id = (int16_t)p.path_front_near_id;
        r.ob_index = -2; r.ob_pathid = 0;
        r.n_traj = (uint16_t)g_calllog_n;
        r.afresh_planning = p.afresh_planning; r.ob_flag = 2; r.acc_flag = p.acc_flag; r.cnt = po.cnt;
        if (path_xy) {
            double* o = path_xy + c * path_xy_stride;
            for (int i = 0; i < 200; ++i) { o[i] = CPlanning::last_Bpoints[i].x; o[200 + i] = CPlanning::last_Bpoints[i].y; }
        }
        if (path_ll) {
            double* o = path_ll + c * path_ll_stride;
            for (int i = 0; i < 100; ++i) { o[i] = po.pnts[i].x; o[100 + i] = po.pnts[i].y; }
        }
        fill_frames(po, g_app.out_status, g_cap_ctrl ? g_cap_ctrl + c * g_cap_stride : nullptr,
                    g_cap_status ? g_cap_status + c * g_cap_stride : nullptr);
        if (g_hook) g_hook(c, g_hook_user);   // may rewrite the inputs of the cycles that follow
    }
    if (carry_out) {
        CDecision& d = CDecision::Instance();
        CPlanning& p = CPlanning::Instance();
        dp_carry& k = *carry_out;
        memset(&k, 0, sizeof(k));
        k.leftlight_time = d.leftlight_time; k.rightlight_time = d.rightlight_time; k.velocity_expect = d.z_velocity_expect;
        k.aim_x = p.aimpoint_far.Aim_point.x; k.aim_y = p.aimpoint_far.Aim_point.y; k.aim_dir = p.aimpoint_far.Aim_point.dir;
        k.aim_id = p.aimpoint_far.Aim_id;
        k.obsavoid_time = ref_obsavoid_time; k.no_obsavoid_time = ref_no_obsaviod_time; k.frontobs_time = ref_frontobs_time;
        k.plan_his_behavior = p.his_behavior;
        k.path_near_id = p.path_near_id;
        k.behavior = d.z_behavior; k.target_roadnum = d.z_target_roadnum; k.target_lanenum = d.z_target_lanenum;
        k.light_status = d.z_light_status; k.behavior_to_dlg = d.z_behavior_to_dlg;
        k.his_behavior = d.his_behavior; k.his_target_lanenum = d.his_target_lanenum; k.his_light_status = d.his_light_status;
        k.lanechg_status = d.z_segment_lanechg_status; k.obsavoid_status = d.z_segment_obsavoid_status;
        k.plan_count = 0;   // 'count' is a local of CPlanningThread (Planning.cpp:51): only PlanningOut.cnt is visible
    }
    if (last_path_out)
        for (int i = 0; i < 200; ++i) { last_path_out[i] = CPlanning::last_Bpoints[i].x; last_path_out[200 + i] = CPlanning::last_Bpoints[i].y; }
    co_drop_all();
    return g_msgbox ? 1 : 0;
}

// ---- the V2X handlers of the unmodified reference, called as SegmentDecision calls them (Decision.cpp:251-283): fresh flags,
// empty path vectors; mode 1 routes status 4 to V2XConstructionEventTemporal (the alternative commented out at :2156) ----
extern "C" int ref_v2x_event(int n, const dp_scene_hdr* hdr, const dp_v2x_data* v2x, const double* wp_lat, const double* wp_lng, int mode,
                             dp_v2x_flags* out) {
    CDecision& d = CDecision::Instance();
    for (int s = 0; s < n; ++s) {
        const dp_scene_hdr& h = hdr[s];
        const dp_v2x_data& v = v2x[s];
        LocationOut lo;
        memset(&lo, 0, sizeof(lo));
        lo.globalpoint.x = h.x; lo.globalpoint.y = h.y; lo.globalpoint.dir = h.dir;
        lo.gpspoint.lat = v.ego_lat; lo.gpspoint.lng = v.ego_lng;
        for (int i = 0; i < LANESUM; ++i) lo.id[i] = h.id[i];
        lo.lane_num = h.lane_num; lo.road_num = h.road_num; lo.pos = (BYTE)h.pos;
        V2X_Data x;
        memset(&x, 0, sizeof(x));
        x.PedesDistance = v.ped_distance; x.PedesLatitude = v.ped_lat; x.PedesLongitude = v.ped_lng; x.PedesDirection = v.ped_direction;
        x.SPATLaneOccupied = v.spat_lane_occupied; x.SPATState = v.spat_state; x.RSILatitude = v.rsi_lat; x.RSILongitude = v.rsi_lng;
        x.V2XWarnStatus = v.warn_status;
        vector<WarningPoint> list;
        for (int k = 0; k < v.wp_count; ++k) { WarningPoint w; w.latitude = wp_lat[v.wp_first + k]; w.longitude = wp_lng[v.wp_first + k]; list.push_back(w); }
        bool pedestrian_flag = false, construction_flag = false;
        WORD light = 0;
        vector<GlobalPoint2D> f, lf, rf;
        if (mode == 1 && v.warn_status == 4) d.V2XConstructionEventTemporal(lo, x, list, f, lf, rf, construction_flag);
        else d.V2XEventDecision(lo, x, list, f, lf, rf, pedestrian_flag, light, construction_flag);
        memset(&out[s], 0, sizeof(out[s]));
        out[s].light_flag = light; out[s].construction_flag = construction_flag; out[s].pedestrian_flag = pedestrian_flag;
        out[s].lng_distance = 9999; out[s].lat_distance = 9999;   // (locals of the handlers: not observable)
    }
    return 0;
}

// Batch form with the [cycle][scene] layout shared with liboracle.so and the CUDA path.
// Scenes run sequentially on ONE thread: the reference is not re-entrant (function statics
// Decision.cpp:915-917, static last_Bpoints Planning.cpp:6).  *seconds covers the whole loop.
extern "C" long long ref_run_batch(int n, int cycles, int max_obs, const dp_scene_hdr* hdr, const double* ox,
                                   const double* oy, dp_plan_record* rec, double* path_xy, double* path_ll,
                                   ref_call* calls, int32_t* n_calls, int calls_cap, dp_carry* carry_out,
                                   double* last_path_out, double* seconds, long long* traj_scored) {
    long long before = g_search_calls;
    int worst = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int s = 0; s < n; ++s) {
        g_cap_ctrl = g_cap_ctrl_base ? g_cap_ctrl_base + s : nullptr;
        g_cap_status = g_cap_status_base ? g_cap_status_base + s : nullptr;
        g_cap_stride = n;
        int rc = ref_run_episode(cycles, hdr + s, n, ox + (size_t)s * max_obs, oy + (size_t)s * max_obs, (long)n * max_obs,
                                 rec + s, n, path_xy ? path_xy + (size_t)s * 400 : nullptr, (long)n * 400,
                                 path_ll ? path_ll + (size_t)s * 200 : nullptr, (long)n * 200,
                                 calls ? calls + (size_t)s * calls_cap : nullptr, n_calls ? n_calls + s : nullptr, n,
                                 (long)n * calls_cap, calls_cap, carry_out ? carry_out + s : nullptr,
                                 last_path_out ? last_path_out + (size_t)s * 400 : nullptr);
        g_cap_ctrl = nullptr; g_cap_status = nullptr;
        if (rc < 0) return rc;
        if (rc > worst) worst = rc;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (traj_scored) *traj_scored = g_search_calls - before;
    return worst;
}

extern "C" long long ref_search_calls() { return g_search_calls; }
extern "C" int ref_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(dp_scene_hdr);
        case 1: return (int)sizeof(dp_plan_record);
        case 2: return (int)sizeof(dp_carry);
        case 3: return (int)sizeof(ref_call);
        case 4: return (int)sizeof(dp_trace_record);
    }
    return -1;
}
