// oracle/v2x_oracle.cpp -- TEST INFRASTRUCTURE (CPU oracle), see v2x_oracle.h.  Quirks of the reference kept on purpose:
//   * V2XPedestrianJudge loads the 200 points from the ego's own index (Decision.cpp:1839), the road-works handlers from
//     index + ID_MORE (:2029, :2211);
//   * lat_distance_last / lat_distance_time are LOCALS (:1851-1852), so the "approaching for 5 frames" test can never pass and a
//     pedestrian between 1.5 m and one and a half lanes to the left is always ignored (:1943-1949); every pedestrian to the
//     RIGHT of the path (negative LatDis) within 100 m ahead raises the flag (:1950);
//   * the nearest road-works point is kept in `float` variables (:2021-2022, :2103-2104): its latitude / longitude are rounded to
//     single precision (about 0.2 m north-south, 0.8 m east-west at the datum) before the lateral test;
//   * that lateral test uses the path index of the LAST list entry, not of the nearest one (:2118-2122);
//   * `min >= d` keeps the LAST of equally distant points and lets a point behind the path start (-9999) win (:2100).
// Out-of-bounds reads of the reference (path[index + 1] when the nearest point is the last one, any access to an empty path) are
// undefined there; here they set `ub` and clear the flags, and the parity tests keep away from them.
#include "v2x_oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace oracle {
namespace {
using spec::P2;

std::vector<P2> load_front(const MapView& m, int gl, int id, int first_off) {
    std::vector<P2> v;
    const int cnt = m.lane_size(gl);
    const int a = std::min(cnt, id + first_off), b = std::min(cnt, id + 200 + first_off);
    for (int i = a; i < b; ++i) v.push_back(m.pt(gl, i));
    return v;
}
// the longitudinal distance idiom of the handlers (Decision.cpp:1899-1922, :2078-2098)
double lng_to(const std::vector<P2>& path, int index, double at_zero) {
    if (index >= 2) {
        double d = 0;
        for (int i = 0; i < index; ++i) {
            const double dx = path[i].x - path[i + 1].x, dy = path[i].y - path[i + 1].y;
            d += std::sqrt(dx * dx + dy * dy);
        }
        return d;
    }
    if (index == 1) {
        const double dx = path[0].x - path[1].x, dy = path[0].y - path[1].y;
        return std::sqrt(dx * dx + dy * dy);
    }
    return at_zero;
}
P2 to_global(const dp_params& p, double lat, double lng) {
    P2 g;
    spec::wgs84_to_global(spec::Datum{p.lat0, p.lng0, p.k_lat, p.k_lng}, lat, lng, &g.x, &g.y);
    return g;
}
struct Points { std::vector<double> lat, lng; };
// Decision.cpp:2036-2049 / :2345-2358; false: "no road-works information"
bool warning_list(const dp_v2x_data& v, const double* wp_lat, const double* wp_lng, Points& l) {
    for (int k = 0; k < v.wp_count; ++k) { l.lat.push_back(wp_lat[v.wp_first + k]); l.lng.push_back(wp_lng[v.wp_first + k]); }
    if (l.lat.empty() && (v.rsi_lat == 0 || v.rsi_lng == 0)) return false;
    if (l.lat.empty()) { l.lat.push_back(v.rsi_lat); l.lng.push_back(v.rsi_lng); }
    return true;
}
struct Nearest { double min_d; float lat, lng; int index; };
// Decision.cpp:2051-2112 / :2360-2405
Nearest nearest_works(const dp_params& p, const std::vector<P2>& path, const Points& l) {
    Nearest r{9999, 0.f, 0.f, 0};
    for (size_t k = 0; k < l.lat.size(); ++k) {
        const P2 g = to_global(p, l.lat[k], l.lng[k]);
        r.index = spec::nearest_id(g, path.data(), (int)path.size());
        const double d = lng_to(path, r.index, -9999);
        if (r.min_d >= d) { r.min_d = d; r.lat = (float)l.lat[k]; r.lng = (float)l.lng[k]; }
    }
    return r;
}

void pedestrian(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const dp_v2x_data& v, dp_v2x_flags* o) {
    const int gl = m.lane_index(h.road_num, h.lane_num);
    const std::vector<P2> path = load_front(m, gl, h.id[h.lane_num - 1], 0);
    const double width = 3.75 + 3.75 / 2;
    const P2 ped = to_global(p, v.ped_lat, v.ped_lng);
    bool flag = false;
    if (v.ped_distance >= 0 && v.ped_distance <= 100) {
        const int index = spec::nearest_id(ped, path.data(), (int)path.size());
        if (index + 1 >= (int)path.size()) { o->ub = 1; return; }
        const double lat = spec::lat_dis(ped, path[index], path[index + 1], p.epsilon);
        const double lng = lng_to(path, index, v.ped_distance > 5 ? -9999.0 : 0.0);
        const int lat_time = (lat < 0) ? 1 : 0;              // lat_distance_last = 0, lat_distance_time = 0 are locals
        o->lng_distance = lng; o->lat_distance = lat;
        if (lng >= 0 && lng <= 100) {
            if (lat >= width) flag = false;
            else if (lat > 1.5 && lat < width) flag = !(lat_time < 5);
            else if (lat <= 1.5 || v.ped_direction == 1) flag = true;
        } else flag = false;
    }
    o->pedestrian_flag = flag;
}

void signal_light(const dp_v2x_data& v, dp_v2x_flags* o) {
    if (v.spat_lane_occupied == 1) {
        if (v.spat_state == 3 || v.spat_state == 7) o->light_flag = 1;
        else if (v.spat_state == 6) o->light_flag = 2;
        else o->light_flag = 0;
    }
}

void road_works(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const dp_v2x_data& v, const double* wp_lat,
                const double* wp_lng, dp_v2x_flags* o) {
    const int gl = m.lane_index(h.road_num, h.lane_num);
    const std::vector<P2> path = load_front(m, gl, h.id[h.lane_num - 1], p.id_more);
    Points l;
    if (!warning_list(v, wp_lat, wp_lng, l)) return;
    if (path.empty()) { o->ub = 1; return; }
    const Nearest r = nearest_works(p, path, l);
    if (r.index + 1 >= (int)path.size()) { o->ub = 1; return; }
    const P2 g = to_global(p, (double)r.lat, (double)r.lng);
    const double nld = spec::lat_dis(g, path[r.index], path[r.index + 1], p.epsilon);
    o->lng_distance = r.min_d; o->lat_distance = nld;
    if (r.min_d >= 0 && r.min_d <= 100) o->construction_flag = (nld >= 0 && nld < 3.75 / 2);
}

double min_dist(const std::vector<P2>& path, P2 q) {
    double best = 9999;
    for (const P2& a : path) { const double d = spec::calc_distance(q, a); if (d < best) best = d; }
    return best;
}

void road_works_temporal(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const dp_v2x_data& v, const double* wp_lat,
                         const double* wp_lng, dp_v2x_flags* o) {
    const int lane = h.lane_num, gl = m.lane_index(h.road_num, lane), id = h.id[lane - 1];
    if (id >= m.lane_size(gl)) { o->ub = 1; return; }         // the reference indexes the lane at Id_CurLane (Decision.cpp:2187)
    const int lane_sum = m.lanes_of(h.road_num), chg = m.attr(gl, id);
    const std::vector<P2> F = load_front(m, gl, id, p.id_more);
    std::vector<P2> LF, RF;
    if (chg == 1 && lane > 1) {
        const int idl = h.id[lane - 2], cl = m.lane_size(gl - 1);
        if (idl > 0 && idl < cl) LF = load_front(m, gl - 1, idl, p.id_more);
    }
    if (chg == 2 && lane < lane_sum) {
        const int idr = h.id[lane], cr = m.lane_size(gl + 1);
        if (idr > 0 && idr < cr) RF = load_front(m, gl + 1, idr, p.id_more);
    }
    const P2 ego = to_global(p, v.ego_lat, v.ego_lng), rsi = to_global(p, v.rsi_lat, v.rsi_lng);
    const double rsi_distance = spec::calc_distance(ego, rsi);
    if (!(rsi_distance >= 0 && rsi_distance <= 100)) return;
    const double f = min_dist(F, rsi), lf = min_dist(LF, rsi), rf = min_dist(RF, rsi);
    if ((lf < rf) && (lf < f) && (lf < 2)) return;
    if ((rf < lf) && (rf < f) && (rf < 2)) return;
    if ((f < lf) && (f < rf) && (f < 2)) {
        Points l;
        if (!warning_list(v, wp_lat, wp_lng, l)) return;
        const Nearest r = nearest_works(p, F, l);
        o->lng_distance = r.min_d;
        o->construction_flag = (r.min_d >= 0 && r.min_d <= 100);
    }
}
}  // namespace

void v2x_event(const MapView& m, const dp_params& p, const dp_scene_hdr& h, const dp_v2x_data& v, const double* wp_lat,
               const double* wp_lng, int mode, dp_v2x_flags* o) {
    std::memset(o, 0, sizeof(*o));
    o->lng_distance = 9999; o->lat_distance = 9999;
    if (v.warn_status == 3) signal_light(v, o);
    else if (v.warn_status == 4) { if (mode == 1) road_works_temporal(m, p, h, v, wp_lat, wp_lng, o); else road_works(m, p, h, v, wp_lat, wp_lng, o); }
    else if (v.warn_status == 5) pedestrian(m, p, h, v, o);
    if (o->ub) { o->light_flag = 0; o->construction_flag = 0; o->pedestrian_flag = 0; o->lng_distance = 9999; o->lat_distance = 9999; }
}
void v2x_apply(const dp_v2x_flags& f, dp_plan_record& r) {
    if (f.pedestrian_flag || f.light_flag == 1) { r.brakespeed = 0.0; r.acc_flag = 1; r.des_acc = -3.0; }
    else if (f.construction_flag && r.brakespeed > 3.0) r.brakespeed = 3.0;
}
}  // namespace oracle
