// oracle/world_spec.cpp -- TEST INFRASTRUCTURE (CPU oracle), see world_spec.h.
#include "world_spec.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace oracle {

void world_default_params(dp_world_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->a_max = 3.0; p->loc_back = 8; p->loc_fwd = 56; p->end_margin = 160;
}

namespace {
inline double seg_len(const MapView& m, int gl, int i, double* ddx, double* ddy) {
    const spec::P2 a = m.pt(gl, i), b = m.pt(gl, i + 1);
    *ddx = b.x - a.x; *ddy = b.y - a.y;
    return std::sqrt(*ddx * *ddx + *ddy * *ddy);
}
}  // namespace

void world_step(const MapView& m, const dp_params& p, const dp_world_params& wp, dp_scene_hdr& h, dp_agent* ag, double* ox,
                double* oy, const dp_plan_record* rec, const double* lx, const double* ly) {
    const int n_obs = h.n_obs;
    const double dt = h.period_ms / 1000.0;
    if (rec) {
        // ---- ego: bounded approach to the planned speed, then a walk along the carried local path ----
        const int gl = m.lane_index(h.road_num, h.lane_num), cnt = m.lane_size(gl);
        const bool at_end = h.id[h.lane_num - 1] >= cnt - wp.end_margin;
        const double v = h.velocity, vt = rec->brakespeed * 3.6, dvm = wp.a_max * 3.6 * dt;
        double dv = vt - v;
        if (dv > dvm) dv = dvm;
        if (dv < -dvm) dv = -dvm;
        double vn = v + dv;
        if (vn < 0) vn = 0;
        if (at_end) vn = 0;
        const double ds = at_end ? 0.0 : (v + vn) * 0.5 / 3.6 * dt;
        double x = h.x, y = h.y, dir = h.dir;
        if (ds > 0) {
            int i = rec->afresh_planning ? 0 : rec->path_near_id;   // a fresh path starts at the ego (Planning.cpp:596-611, :845-877)
            if (i < 0) i = 0;
            if (i > DP_PATH_POINTS - 2) i = DP_PATH_POINTS - 2;
            double rem = ds;
            for (;;) {
                const double ax = lx[i], ay = ly[i], bx = lx[i + 1], by = ly[i + 1];
                const double ddx = bx - ax, ddy = by - ay;
                const double L = std::sqrt(ddx * ddx + ddy * ddy);
                if (rem < L || i == DP_PATH_POINTS - 2) {
                    if (L > 0) {
                        double t = rem / L;
                        if (t > 1) t = 1;
                        x = ax + t * ddx; y = ay + t * ddy;
                        dir = spec::calc_global_dir(spec::P2{ax, ay}, spec::P2{bx, by}, p.epsilon, p.pi);
                    } else { x = ax; y = ay; }
                    break;
                }
                rem -= L; ++i;
            }
        }
        h.x = x; h.y = y; h.dir = dir; h.velocity = vn;
        // ---- agents: constant speed along their lanes ----
        for (int k = 0; k < n_obs; ++k) {
            dp_agent& a = ag[k];
            const int cntk = m.lane_size(a.lane);
            double u = a.u + a.v * dt, ddx, ddy;
            int i = a.i;
            while (i < cntk - 2) {
                const double L = seg_len(m, a.lane, i, &ddx, &ddy);
                if (u < L) break;
                u -= L; ++i;
            }
            if (i >= cntk - 2) {
                i = cntk - 2;
                const double L = seg_len(m, a.lane, i, &ddx, &ddy);
                if (u > L) u = L;
            }
            a.u = u; a.i = i;
        }
    }
    // ---- obstacle points of the agents ----
    for (int k = 0; k < n_obs; ++k) {
        const dp_agent& a = ag[k];
        double ddx, ddy;
        const double L = seg_len(m, a.lane, a.i, &ddx, &ddy);
        const spec::P2 q = m.pt(a.lane, a.i);
        if (L > 0) {
            const double t = a.u / L;
            const double px = q.x + t * ddx, py = q.y + t * ddy;
            ox[k] = px + a.lat * (-(ddy / L));
            oy[k] = py + a.lat * (ddx / L);
        } else { ox[k] = q.x; oy[k] = q.y; }
    }
    // ---- localisation: nearest point per lane inside a window around the previous index, nearest lane ----
    const int nl = m.lanes_of(h.road_num);
    double best = std::numeric_limits<double>::infinity();
    int bl = 0;
    for (int l = 0; l < nl && l < DP_LANESUM; ++l) {
        const int gl = m.lane_index(h.road_num, l + 1), cnt = m.lane_size(gl);
        int lo = h.id[l] - wp.loc_back, hi = h.id[l] + wp.loc_fwd;
        if (lo < 0) lo = 0;
        if (hi > cnt - 1) hi = cnt - 1;
        double be = std::numeric_limits<double>::infinity();
        int bi = lo;
        for (int i = lo; i <= hi; ++i) {
            const spec::P2 q = m.pt(gl, i);
            const double dx = h.x - q.x, dy = h.y - q.y;
            const double e = dx * dx + dy * dy;
            if (e < be) { be = e; bi = i; }
        }
        h.id[l] = bi;
        if (be < best) { best = be; bl = l; }
    }
    h.lane_num = (uint16_t)(bl + 1);
}

void pack_frames(const dp_params& p, const dp_plan_record& r, const double* px, const double* py, dp_ctrl_frame* ctrl,
                 dp_status_frame* status) {
    if (ctrl) {
        std::memset(ctrl, 0, sizeof(*ctrl));
        ctrl->cnt = r.cnt; ctrl->apa = 0; ctrl->brakedis = r.mindist_lon; ctrl->brake_speed = 0; ctrl->desacc_vd = r.acc_flag;
        ctrl->desacc = r.des_acc; ctrl->desspd = r.brakespeed; ctrl->desstr = 0; ctrl->desstr_vd = 0; ctrl->light = r.light;
        ctrl->radius = r.radius; ctrl->road_type = 0; ctrl->sstop = 1;
        const spec::Datum dm{p.lat0, p.lng0, p.k_lat, p.k_lng};
        for (int i = 0; i < DP_OUT_POINTS; ++i) spec::global_to_wgs84(dm, px[2 * i], py[2 * i], &ctrl->pnts[i][0], &ctrl->pnts[i][1]);
    }
    if (status) {
        std::memset(status, 0, sizeof(*status));
        status->afresh_cause = r.afresh_cause; status->near_ob_dist = r.mindist_lon; status->planspeed = r.brakespeed;
        status->planacc = r.des_acc; status->trafficlight = r.light;
        for (int i = 0; i < DP_OUT_POINTS; ++i) { status->path_points[i][0] = px[2 * i]; status->path_points[i][1] = py[2 * i]; }
    }
}

}  // namespace oracle
