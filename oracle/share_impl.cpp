// oracle/share_impl.cpp -- TEST INFRASTRUCTURE. CShare member bodies for the reference build:
// thin adapters from the reference's call signatures (compat/Share.h) to the frozen
// specification in cshare_spec.cpp.
#include "compat/Share.h"
#undef min
#undef max
#include "cshare_spec.h"
#include "ref_api.h"

namespace { spec::Datum g_datum{23.0, 113.0, 1.0 / 110574.0, 1.0 / 102470.0}; }
extern "C" void share_set_datum(double lat0, double lng0, double k_lat, double k_lng) {
    g_datum = spec::Datum{lat0, lng0, k_lat, k_lng};
}
extern "C" {
long long g_search_calls = 0;   // trajectories scored (one per SearchObstacle call)
ref_call* g_calllog = nullptr;  // per-cycle call log, owned by the harness
int g_calllog_n = 0, g_calllog_cap = 0;
}

bool CShare::SearchObstacle(vector<GlobalPoint2D> path, vector<ObPoint> obs, double lat_min, double lat_max,
                            double& dis_lat, double& dis_lng, ObPoint& ob, WORD& pathid) {
    static_assert(sizeof(GlobalPoint2D) == sizeof(spec::P2), "layout");
    vector<spec::P2> o(obs.size());
    for (size_t i = 0; i < obs.size(); ++i) o[i] = spec::P2{obs[i].x, obs[i].y};
    ++g_search_calls;
    spec::SearchResult r = spec::search_obstacle(reinterpret_cast<const spec::P2*>(path.data()), (int)path.size(),
                                                 o.data(), (int)o.size(), lat_min, lat_max);
    if (g_calllog && g_calllog_n < g_calllog_cap) {
        ref_call& c = g_calllog[g_calllog_n];
        c.lat_min = lat_min; c.lat_max = lat_max; c.dis_lat = r.dis_lat; c.dis_lng = r.dis_lng;
        c.n_path = (int32_t)path.size(); c.ob_index = (int16_t)r.ob_index; c.pathid = (uint16_t)r.pathid;
        c.found = r.found;
    }
    ++g_calllog_n;
    dis_lat = r.dis_lat;
    dis_lng = r.dis_lng;
    pathid = (WORD)r.pathid;
    if (r.found) ob = obs[r.ob_index];
    return r.found;
}
vector<GlobalPoint2D> CShare::CreateNewPath(vector<GlobalPoint2D> path, double offset) {
    vector<GlobalPoint2D> out(path.size());
    spec::create_new_path(reinterpret_cast<const spec::P2*>(path.data()), (int)path.size(), offset,
                          reinterpret_cast<spec::P2*>(out.data()));
    return out;
}
void CShare::BezierPlanning(GlobalPoint3D s, GlobalPoint3D a, GlobalPoint2D out[], int n) {
    spec::bezier_planning(spec::P3{s.x, s.y, s.dir}, spec::P3{a.x, a.y, a.dir}, reinterpret_cast<spec::P2*>(out), n);
}
void CShare::MeanPoints(GlobalPoint2D in[], int n_in, GlobalPoint2D out[], int n_out) {
    spec::mean_points(reinterpret_cast<const spec::P2*>(in), n_in, reinterpret_cast<spec::P2*>(out), n_out);
}
double CShare::CalcDistance(GlobalPoint2D a, GlobalPoint2D b) { return spec::calc_distance({a.x, a.y}, {b.x, b.y}); }
double CShare::CalcDistance(GPSPoint2D a, GPSPoint2D b) {
    double ax, ay, bx, by;
    spec::wgs84_to_global(g_datum, a.lat, a.lng, &ax, &ay);
    spec::wgs84_to_global(g_datum, b.lat, b.lng, &bx, &by);
    return spec::calc_distance({ax, ay}, {bx, by});
}
double CShare::CalcGlobalDir(GlobalPoint2D a, GlobalPoint2D b) {
    return spec::calc_global_dir({a.x, a.y}, {b.x, b.y}, EPSILON, PI);
}
int CShare::NearestId(GlobalPoint2D q, vector<GlobalPoint2D> pts) {
    return spec::nearest_id({q.x, q.y}, reinterpret_cast<const spec::P2*>(pts.data()), (int)pts.size());
}
double CShare::LatDis(GlobalPoint2D q, GlobalPoint2D pt, GlobalPoint2D nx) {
    return spec::lat_dis({q.x, q.y}, {pt.x, pt.y}, {nx.x, nx.y}, EPSILON);
}
GlobalPoint2D CShare::WGS84ToGlobal(GPSPoint2D g) {
    GlobalPoint2D p;
    spec::wgs84_to_global(g_datum, g.lat, g.lng, &p.x, &p.y);
    return p;
}
GPSPoint2D CShare::GlobalToWGS84(GlobalPoint2D p) {
    GPSPoint2D g;
    spec::global_to_wgs84(g_datum, p.x, p.y, &g.lat, &g.lng);
    return g;
}
