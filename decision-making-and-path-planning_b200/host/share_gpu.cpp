// host/share_gpu.cpp -- CShare members for host/Share.h: the heavy operators (A1-A4) are CUDA launches
// through the C ABI; failure to reach the GPU aborts (there is no CPU path to fall back to).
#include "Share.h"
#undef min
#undef max
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace {
dp_ctx* g_ctx = nullptr;
dp_params g_params;
long long g_calls = 0;
void must(int rc, const char* what) {
    if (rc != DP_OK) { fprintf(stderr, "libdmpp_b200: %s failed (%d): %s\n", what, rc, dp_last_error()); abort(); }
}
}  // namespace

dp_ctx* CShare::Context() {
    if (!g_ctx) {
        dp_default_params(&g_params);
        must(dp_create(&g_ctx, 0, &g_params, 64, 64), "dp_create");
    }
    return g_ctx;
}
long long CShare::SearchCalls() { return g_calls; }

bool CShare::SearchObstacle(vector<GlobalPoint2D> path, vector<ObPoint> obs, double lat_min, double lat_max, double& dis_lat,
                            double& dis_lng, ObPoint& ob, WORD& pathid) {
    const int P = (int)path.size(), N = (int)obs.size();
    vector<double> px(P), py(P), ox(N), oy(N);
    for (int i = 0; i < P; ++i) { px[i] = path[i].x; py[i] = path[i].y; }
    for (int i = 0; i < N; ++i) { ox[i] = obs[i].x; oy[i] = obs[i].y; }
    const int32_t off[2] = {0, P};
    dp_search_slot s;
    must(dp_search_obstacle(Context(), 1, off, px.data(), py.data(), ox.data(), oy.data(), N, &lat_min, &lat_max, &s), "dp_search_obstacle");
    ++g_calls;
    dis_lat = s.dis_lat; dis_lng = s.dis_lng; pathid = s.pathid;
    if (s.found) ob = obs[s.ob_index];
    return s.found != 0;
}
vector<GlobalPoint2D> CShare::CreateNewPath(vector<GlobalPoint2D> path, double offset) {
    const int P = (int)path.size();
    vector<GlobalPoint2D> out(P);
    if (P == 0) return out;
    vector<double> px(P), py(P), ox(P), oy(P);
    for (int i = 0; i < P; ++i) { px[i] = path[i].x; py[i] = path[i].y; }
    const int32_t off[2] = {0, P};
    must(dp_create_new_path(Context(), 1, off, px.data(), py.data(), &offset, ox.data(), oy.data()), "dp_create_new_path");
    for (int i = 0; i < P; ++i) { out[i].x = ox[i]; out[i].y = oy[i]; }
    return out;
}
void CShare::BezierPlanning(GlobalPoint3D s, GlobalPoint3D a, GlobalPoint2D out[], int n) {
    if (n != DP_PATH_POINTS) { fprintf(stderr, "BezierPlanning: n must be %d\n", DP_PATH_POINTS); abort(); }
    const double poses[6] = {s.x, s.y, s.dir, a.x, a.y, a.dir};
    vector<double> o(2 * DP_PATH_POINTS);
    must(dp_bezier_planning(Context(), 1, poses, o.data()), "dp_bezier_planning");
    for (int i = 0; i < n; ++i) { out[i].x = o[i]; out[i].y = o[DP_PATH_POINTS + i]; }
}
void CShare::MeanPoints(GlobalPoint2D in[], int n_in, GlobalPoint2D out[], int n_out) {
    if (n_out != DP_PATH_POINTS) { fprintf(stderr, "MeanPoints: n_out must be %d\n", DP_PATH_POINTS); abort(); }
    vector<double> px(n_in > 0 ? n_in : 1), py(n_in > 0 ? n_in : 1), o(2 * DP_PATH_POINTS);
    for (int i = 0; i < n_in; ++i) { px[i] = in[i].x; py[i] = in[i].y; }
    const int32_t off[2] = {0, n_in > 0 ? n_in : 0};
    must(dp_mean_points(Context(), 1, off, px.data(), py.data(), o.data()), "dp_mean_points");
    for (int i = 0; i < n_out; ++i) { out[i].x = o[i]; out[i].y = o[DP_PATH_POINTS + i]; }
}

// ---- scalar helpers: same definitions as the device code (csrc/dp_device.cuh) ----
static double h_atan(double z) {
    const double PI_2 = 1.57079632679489661923, PI_4 = 0.78539816339744830962;
    const bool neg = z < 0;
    double a = neg ? -z : z;
    const bool inv = a > 1.0;
    if (inv) a = 1.0 / a;
    const bool shift = a > 0.41421356237309503;
    const double w = shift ? (a - 1.0) / (a + 1.0) : a, w2 = w * w;
    double p = 1.0 / 47.0;
    for (int n = 22; n >= 0; --n) p = std::fma(-w2, p, 1.0 / (double)(2 * n + 1));
    double r = w * p;
    if (shift) r = PI_4 + r;
    if (inv) r = PI_2 - r;
    return neg ? -r : r;
}
double CShare::CalcDistance(GlobalPoint2D a, GlobalPoint2D b) {
    const double dx = a.x - b.x, dy = a.y - b.y;
    return std::sqrt(dx * dx + dy * dy);
}
double CShare::CalcDistance(GPSPoint2D a, GPSPoint2D b) { return CalcDistance(WGS84ToGlobal(a), WGS84ToGlobal(b)); }
double CShare::CalcGlobalDir(GlobalPoint2D a, GlobalPoint2D b) {
    double angle;
    if (std::fabs(b.x - a.x) < EPSILON && std::fabs(b.y - a.y) < EPSILON) angle = 0;
    else if (std::fabs(b.x - a.x) < EPSILON) angle = (b.y > a.y) ? PI / 2 : 3 * PI / 2;
    else {
        angle = h_atan((b.y - a.y) / (b.x - a.x));
        if (b.x < a.x) angle = angle + PI;
        else if ((b.x > a.x) && (b.y < a.y)) angle = angle + 2 * PI;
    }
    return angle * 180 / PI;
}
int CShare::NearestId(GlobalPoint2D q, vector<GlobalPoint2D> pts) {
    double best = 9999; int id = 0;
    for (size_t i = 0; i < pts.size(); ++i) { const double d = CalcDistance(q, pts[i]); if (d < best) { best = d; id = (int)i; } }
    return id;
}
double CShare::LatDis(GlobalPoint2D c, GlobalPoint2D pt, GlobalPoint2D nx) {
    double l;
    if (std::fabs(pt.x - nx.x) > EPSILON) {
        const double k = (pt.y - nx.y) / (pt.x - nx.x);
        l = std::fabs((c.y - pt.y) - k * (c.x - pt.x)) / std::sqrt(1 + k * k);
    } else l = std::fabs(pt.x - c.x);
    if (l < EPSILON) return 0;
    const double s = (nx.x - pt.x) * (c.y - pt.y) - (nx.y - pt.y) * (c.x - pt.x);
    return l * (s > 0 ? 1 : -1);
}
GlobalPoint2D CShare::WGS84ToGlobal(GPSPoint2D g) {
    Context();
    GlobalPoint2D p;
    p.y = (g.lat - g_params.lat0) / g_params.k_lat;
    p.x = (g.lng - g_params.lng0) / g_params.k_lng;
    return p;
}
GPSPoint2D CShare::GlobalToWGS84(GlobalPoint2D p) {
    Context();
    GPSPoint2D g;
    g.lat = std::fma(p.y, g_params.k_lat, g_params.lat0);
    g.lng = std::fma(p.x, g_params.k_lng, g_params.lng0);
    return g;
}
