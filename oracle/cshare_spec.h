// oracle/cshare_spec.h -- TEST INFRASTRUCTURE (CPU oracle). Never linked into the product.
//
// Double-precision SPECIFICATION of the geometry operators the reference calls through its
// `CShare` base class (`Share.h` / `Share.cpp` of the un-shipped "GAC_Auotpilot_DP" MFC project).
// The reference at /root/reference does NOT contain their source, and has no tests or golden
// vectors for them, so every function below is PARITY UNPINNED against the original: the
// signatures, sign conventions and sentinels are recovered from the call sites cited on each
// function; the arithmetic is this repo's frozen definition.
//
// Arithmetic discipline (what makes CPU == GPU bit for bit):
//   * only IEEE-754 binary64 +, -, *, /, sqrt, and fma where the text below says "fma";
//   * no implicit contraction: this file is compiled with -ffp-contract=off, the CUDA side with
//     -fmad=false, and both call fma() explicitly at the same places;
//   * sums that the text calls "sequential" are evaluated left to right in index order;
//   * no libm transcendental is used: headings enter through spec_sincos_deg(), a fixed
//     polynomial evaluated with fma, identical on both sides.
#pragma once
#include <cstddef>
#include <cstdint>

namespace spec {

struct P2 { double x, y; };
struct P3 { double x, y, dir; };

// Sentinel written to dis_lat / dis_lng when no obstacle lies in the corridor.  Callers in the
// reference test dis_lng against 13/15/25 without looking at the returned flag
// (Decision.cpp:373,458,922,1149) and Planning pre-loads 999 (Planning.cpp:161-162).
constexpr double NOT_FOUND = 999.0;

struct SearchResult {
    bool   found;     // return value of CShare::SearchObstacle
    double dis_lat;   // signed lateral offset of the selected obstacle, RIGHT of path positive
    double dis_lng;   // arclength from path[0] to path[pathid] (sequential sum of segment lengths)
    int    ob_index;  // index of the selected obstacle in obs[] (-1 when !found)
    int    pathid;    // index of the path point nearest to the selected obstacle (0 when !found)
};

// sqrt(dx*dx + dy*dy) with dx = a.x-b.x, dy = a.y-b.y; two products, one add, one sqrt, no fma.
// Mirrors the in-tree idiom sqrt(pow(dx,2)+pow(dy,2)) (Planning.cpp:413,642,670).
// Call sites: Planning.cpp:509,545; Decision.cpp:1186..1530 (CShare::CalcDistance).
double calc_distance(P2 a, P2 b);

// Heading of a->b in degrees, 0 = east, CCW positive, [0,360).  Same branch structure as the
// in-tree twin CPlanning::GetRoadAngle (Planning.cpp:719-750) but atan is replaced by
// spec_atan() so the value is reproducible.  Call sites: Planning.cpp:519-571 (CalcGlobalDir).
double calc_global_dir(P2 a, P2 b, double epsilon, double pi);

// Index of the point of pts[0..n) nearest to q; strict '<' so the lowest index wins ties
// (loop shape of Planning.cpp:640-648).  Call sites: Decision.cpp:1889,2074,2383 (NearestId).
int nearest_id(P2 q, const P2* pts, int n);

// Signed distance of q from the line pt->pt_next, LEFT positive: operation-for-operation the
// in-tree CPlanning::GetLatDis (Planning.cpp:686-709).  Call sites: Decision.cpp:1895,2118.
double lat_dis(P2 q, P2 pt, P2 pt_next, double epsilon);

// cos/sin of an angle given in degrees.  k = rint(a/90); r = fma(-90,k,a); x = r*(pi/180);
// Taylor polynomials (sin: degree 17, cos: degree 16) in Horner form with fma; quadrant rotation
// by k mod 4.  |error| < 3e-16.
void spec_sincos_deg(double a_deg, double* c, double* s);

// atan(z) for any finite z: argument reduction |z|>1 -> pi/2 - atan(1/|z|), then
// z in [0,1] split at tan(pi/8) by (z-1)/(z+1)+pi/4, odd minimax-free Taylor/Euler series in
// Horner form with fma.  |error| < 1e-15.  Only used by calc_global_dir.
double spec_atan(double z);

// ---------------------------------------------------------------------------------------------
// CShare::SearchObstacle(path, obs, lat_min, lat_max, &dis_lat, &dis_lng, &ob, &pathid)
// Call sites: Planning.cpp:168; Decision.cpp:370,455,811,817,823,830,836,842,943,962.
//
// For every obstacle point o (index order):
//   1. j*  = argmin_j (o.x-p_j.x)^2 + (o.y-p_j.y)^2, value = fma(dx,dx,dy*dy), strict '<'
//            (lowest j wins ties);
//   2. k   = (j* == P-1) ? P-2 : j*;  s = p_{k+1} - p_k;  len = sqrt(fma(s.x,s.x,s.y*s.y));
//   3. gate: if j* == 0     require (o-p_0)    . s >= 0   (obstacle not behind the start)
//            if j* == P-1   require (o-p_{P-1}). s <= 0   (obstacle not beyond the end)
//            where a.b = fma(a.x,b.x,a.y*b.y);
//   4. d   = fma(o.x-p_k.x, s.y, -((o.y-p_k.y)*s.x)) / len   (RIGHT of path positive; len==0 -> d=0)
//   5. in corridor iff lat_min <= d <= lat_max.
// Among in-corridor obstacles the one with the smallest j* is selected (ties: lowest obstacle
// index).  dis_lng = sum_{i<j*} |p_{i+1}-p_i| (sequential, each term sqrt(fma(dx,dx,dy*dy))).
// P < 2 or no obstacle in corridor: found=false, dis_lat=dis_lng=NOT_FOUND, pathid=0.
// ---------------------------------------------------------------------------------------------
SearchResult search_obstacle(const P2* path, int P, const P2* obs, int N,
                             double lat_min, double lat_max);

// Generalisation of search_obstacle to predicted obstacle tracks (BASELINE config 3/5; the
// reference itself has no prediction): obstacle o is at
// ( fma(j, dv[o].x, obs[o].x), fma(j, dv[o].y, obs[o].y) ) when the ego reaches path point j.
// Step 1 compares |o(j) - p_j|^2 over j; steps 2-5 use the obstacle position at j = j*.
// dv == nullptr (or all zero) is exactly search_obstacle.
SearchResult search_obstacle_tracks(const P2* path, int P, const P2* obs, const P2* dv, int N,
                                    double lat_min, double lat_max);

// Predicted agent tracks of BASELINE config 5 (the reference has no prediction: Decision.cpp:162-163 comments the dynamic
// obstacle getters out).  Constant-turn-rate rollout, T positions per agent, every operation pinned:
//   (c, s) = spec_sincos_deg(dtheta);  x_0 = x0, y_0 = y0, v_0 = (vx, vy)  [metres per step];
//   x_{j+1} = x_j + vx_j;  y_{j+1} = y_j + vy_j;  vx_{j+1} = fma(c, vx_j, -(s * vy_j));  vy_{j+1} = fma(s, vx_j, c * vy_j).
// out_x / out_y are written with stride `stride` (the [T x N] track tile of a scene: element j of agent o at j * N + o).
void rollout_ctr(double x0, double y0, double vx, double vy, double dtheta_deg, int T, double* out_x, double* out_y, int stride);

// search_obstacle against a [T x N] track tile: agent o is at (tile_x[j' * N + o], tile_y[j' * N + o]), j' = min(j, T-1), when
// the ego reaches path point j (an agent keeps its last predicted position beyond the horizon).  Step 1 compares
// |o(j) - p_j|^2 over j; steps 2-5 use the position at j = j*.  T == 1 is exactly search_obstacle.
SearchResult search_obstacle_tile(const P2* path, int P, const double* tile_x, const double* tile_y, int T, int N,
                                  double lat_min, double lat_max);

// CShare::CreateNewPath(path, d): lateral-offset copy, d > 0 shifts to the RIGHT of the direction
// of travel (Decision.cpp:629 uses -W for the left lane, :942 -0.3*i for "left avoid").
//   segment for point j: (j, j+1), last point reuses (P-2, P-1);
//   len = sqrt(fma(s.x,s.x,s.y*s.y)); n = (s.y/len, -s.x/len)  (len==0 -> n=(0,0));
//   out_j = ( fma(d,n.x,p_j.x), fma(d,n.y,p_j.y) ).   P < 2: plain copy.
void create_new_path(const P2* path, int P, double d, P2* out);

// CShare::BezierPlanning(start, aim, out, n): cubic Bezier (Planning.cpp:606,863).
//   D = sqrt(fma(ex,ex,ey*ey)), e = aim-start;  L = D/3;
//   P1 = start + L*(cos,sin)(start.dir), P2 = aim - L*(cos,sin)(aim.dir)  (each coord one fma);
//   t_i = i/(n-1), u = 1-t;  b0=u*u*u, b1=3*(u*u)*t, b2=3*u*(t*t), b3=t*t*t (left to right);
//   x_i = fma(b3,x3, fma(b2,x2, fma(b1,x1, b0*x0))), same for y.
void bezier_planning(P3 start, P3 aim, P2* out, int n);

// CShare::MeanPoints(in, n_in, out, n_out): uniform arclength resample (Planning.cpp:872).
//   n_in <= 0: out = 0;  n_in == 1: out = in[0] repeated.
//   cum[0]=0, cum[i+1]=cum[i]+|in[i+1]-in[i]| (sequential);  step = cum[n_in-1]/(n_out-1);
//   for k: s = k*step;  i = largest index <= n_in-2 with cum[i] <= s;
//          seg = cum[i+1]-cum[i];  t = seg > 0 ? (s-cum[i])/seg : 0;
//          out_k = ( fma(t, in[i+1].x-in[i].x, in[i].x), same y );  out_{n_out-1} = in[n_in-1].
void mean_points(const P2* in, int n_in, P2* out, int n_out);

// CShare::GlobalToWGS84 / WGS84ToGlobal (Planning.cpp:209; Decision.cpp:1881,2072,2112,2260):
// equirectangular datum about (lat0,lng0): lat = fma(y,k_lat,lat0), lng = fma(x,k_lng,lng0).
struct Datum { double lat0, lng0, k_lat, k_lng; };
void global_to_wgs84(const Datum& d, double x, double y, double* lat, double* lng);
void wgs84_to_global(const Datum& d, double lat, double lng, double* x, double* y);

}  // namespace spec
