// csrc/dp_device.cuh -- device-side geometry operators (the CShare seam of the reference,
// SURVEY.md 8a rows A1-A5) for sm_100a.  One WARP cooperates on one path/trajectory:
//   * lanes are (obstacle, point-chunk) work items for the nearest-point search, so that
//     N = 10 obstacles still fill 30 of 32 lanes; N >= 32 runs in groups of 32 obstacles;
//   * path points are staged per warp in shared memory in tiles of DP_TILE points (gathered from
//     the L2-resident map with coalesced loads, lateral offset fused into the gather), then read
//     as 16-byte broadcast loads in the inner loop -- candidate paths never touch HBM;
//   * selection is a packed (path index << 16 | obstacle index) warp min-reduction (redux.sync),
//     i.e. nearest-along-path first, lowest obstacle index on ties;
//   * sums the reference evaluates sequentially (arclength) stay sequential: terms are produced
//     in parallel into shared memory, then added in index order, so results are bit-identical
//     to a scalar CPU evaluation.
// Arithmetic: IEEE binary64, compiled with -fmad=false; fma() appears exactly where the
// operator specification says so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dmpp_b200.h"

#define DP_FULL 0xffffffffu
#define DP_TILE 128            // path points staged per warp per tile
#define DP_SCR 256             // per-warp scratch doubles
#define DP_WARPS_PER_BLOCK 4

struct DevMap {
    const double* x; const double* y; const double* dir;
    const double* nx; const double* ny;       // unit RIGHT normal of segment i -> i+1 (0 at a lane's last point)
    const double* lenp;                        // |p[i+1]-p[i]| as sqrt(dx*dx+dy*dy)       (CalcDistance idiom)
    const uint16_t* width; const uint16_t* attr;
    const int32_t* road_lane_base; const int32_t* lane_pt_off;
    const dp_connector* conn;
    int n_roads, n_lanes, n_conn;
};

struct WarpSmem {
    double2 tile[DP_TILE + 1];                 // current path tile (+1 so segments can be formed)
    double scr[DP_SCR];                        // sequential-sum terms / cum[] of MeanPoints
    double2 plan[DP_PATH_POINTS];              // road_points of this cycle (Planning.cpp:115)
};

// A path is never materialised in global memory: it is a recipe evaluated into the tile.
struct PathSrc {
    int kind;          // 0: up to two map segments, 1: window of sm.plan, 2: caller polyline in global memory
    int P;             // number of points
    int base0, step0, n0;   // kind 0: map point index base0 + step0*j for j < n0 (step0 = +1 / -1)
    int base1;              // kind 0: second (forward) segment, points j >= n0
    int s0;                 // kind 1: first plan index
    double d;               // lateral offset, RIGHT positive; kind 0 single segment or kind 2
    const double* gx; const double* gy;   // kind 2
};

struct SearchRes { bool found; double dis_lat, dis_lng; int ob, pathid; };

__device__ __forceinline__ double dp_sq2(double dx, double dy) { return fma(dx, dx, dy * dy); }
__device__ __forceinline__ double dp_dist_plain(double ax, double ay, double bx, double by) {
    double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

// unit right normal of segment a->b per the CreateNewPath specification
__device__ __forceinline__ double2 dp_normal(double2 a, double2 b) {
    double sx = b.x - a.x, sy = b.y - a.y;
    double len = sqrt(dp_sq2(sx, sy));
    if (len > 0) return make_double2(sy / len, -sx / len);
    return make_double2(0.0, 0.0);
}

__device__ __forceinline__ double2 dp_path_point(const DevMap& m, const WarpSmem& sm, const PathSrc& s, int j) {
    if (s.kind == 1) return sm.plan[s.s0 + j];
    if (s.kind == 2) {
        double x = s.gx[j], y = s.gy[j];
        if (s.d != 0.0 && s.P >= 2) {
            int k = (j == s.P - 1) ? s.P - 2 : j;
            double2 n = dp_normal(make_double2(s.gx[k], s.gy[k]), make_double2(s.gx[k + 1], s.gy[k + 1]));
            x = fma(s.d, n.x, x); y = fma(s.d, n.y, y);
        }
        return make_double2(x, y);
    }
    if (j < s.n0) {
        int idx = s.base0 + s.step0 * j;
        double x = m.x[idx], y = m.y[idx];
        if (s.d != 0.0 && s.P >= 2) {
            int jj = min(j, s.P - 2);
            int ni = (s.step0 > 0) ? s.base0 + jj : s.base0 - jj - 1;
            double nx = m.nx[ni], ny = m.ny[ni];
            if (s.step0 < 0) { nx = -nx; ny = -ny; }
            x = fma(s.d, nx, x); y = fma(s.d, ny, y);
        }
        return make_double2(x, y);
    }
    int idx = s.base1 + (j - s.n0);
    return make_double2(m.x[idx], m.y[idx]);
}

// sequential (index-order) sum of sm.scr[0..n): every lane runs the same chain, so the result is
// warp-uniform without a shuffle and costs one issue slot per term.
__device__ __forceinline__ double dp_seq_sum(const WarpSmem& sm, int n, double acc) {
    int j = 0;
    for (; j + 4 <= n; j += 4) {
        double a = sm.scr[j], b = sm.scr[j + 1], c = sm.scr[j + 2], d = sm.scr[j + 3];
        acc += a; acc += b; acc += c; acc += d;
    }
    for (; j < n; ++j) acc += sm.scr[j];
    return acc;
}

// CShare::SearchObstacle (Planning.cpp:168; Decision.cpp:370,455,811-842,943,962), one warp.
// ox/oy: the scene's obstacle arrays in global memory.  Returns a warp-uniform result.
static __device__ SearchRes dp_search_path(const DevMap& m, WarpSmem& sm, const PathSrc& s, const double* __restrict__ ox,
                                    const double* __restrict__ oy, int N, double lo, double hi, int lane) {
    SearchRes r;
    r.found = false; r.dis_lat = DP_NOT_FOUND; r.dis_lng = DP_NOT_FOUND; r.ob = -1; r.pathid = 0;
    const int P = s.P;
    if (P < 2 || N <= 0) return r;
    const int nchunk = (N >= 32) ? 1 : 32 / N;
    const int CS = (P + nchunk - 1) / nchunk;
    const int ngroups = (nchunk == 1) ? (N + 31) / 32 : 1;
    const int tile_n = (s.kind == 1) ? P : DP_TILE;
    unsigned bestkey = 0xffffffffu;
    double bestd = 0.0;
    for (int g = 0; g < ngroups; ++g) {
        const int o = (nchunk == 1) ? g * 32 + lane : lane % N;
        const int c = (nchunk == 1) ? 0 : lane / N;
        const bool active = (nchunk == 1) ? (o < N) : (lane < N * nchunk);
        const double mx = active ? ox[o] : 0.0, my = active ? oy[o] : 0.0;
        const int jlo = c * CS, jhi = min(P, jlo + CS);
        double bd = __longlong_as_double(0x7ff0000000000000LL);
        int bj = jlo;
        for (int t0 = 0; t0 < P; t0 += tile_n) {
            const int tn = min(tile_n, P - t0);
            const double2* pts;
            if (s.kind == 1) pts = sm.plan + s.s0;
            else {
                for (int j = lane; j < tn; j += 32) sm.tile[j] = dp_path_point(m, sm, s, t0 + j);
                __syncwarp();
                pts = sm.tile;
            }
            const int a = max(jlo, t0), b = active ? min(jhi, t0 + tn) : 0;
#pragma unroll 4
            for (int j = a; j < b; ++j) {
                const double2 p = pts[j - t0];
                const double dx = mx - p.x, dy = my - p.y;
                const double d2 = fma(dx, dx, dy * dy);
                if (d2 < bd) { bd = d2; bj = j; }
            }
            __syncwarp();
        }
        for (int cc = 1; cc < nchunk; ++cc) {            // chunks are index-ordered: strict '<' keeps the lowest j
            const int src = (lane % N + cc * N) & 31;
            const double od = __shfl_sync(DP_FULL, bd, src);
            const int oj = __shfl_sync(DP_FULL, bj, src);
            if (lane < N && od < bd) { bd = od; bj = oj; }
        }
        const bool owner = (nchunk == 1) ? active : (lane < N);
        if (owner) {
            const int k = (bj == P - 1) ? P - 2 : bj;
            const double2 pk = dp_path_point(m, sm, s, k), pk1 = dp_path_point(m, sm, s, k + 1);
            const double sx = pk1.x - pk.x, sy = pk1.y - pk.y;
            bool pass = true;
            if (bj == 0) pass = fma(mx - pk.x, sx, (my - pk.y) * sy) >= 0.0;
            else if (bj == P - 1) pass = fma(mx - pk1.x, sx, (my - pk1.y) * sy) <= 0.0;
            const double len = sqrt(dp_sq2(sx, sy));
            double d = 0.0;
            if (len > 0) d = fma(mx - pk.x, sy, -((my - pk.y) * sx)) / len;
            pass = pass && (d >= lo && d <= hi);
            const unsigned key = pass ? (((unsigned)bj << 16) | (unsigned)o) : 0xffffffffu;
            if (key < bestkey) { bestkey = key; bestd = d; }
        }
    }
    const unsigned gmin = __reduce_min_sync(DP_FULL, bestkey);
    if (gmin == 0xffffffffu) return r;
    const int jstar = (int)(gmin >> 16), ostar = (int)(gmin & 0xffffu);
    r.found = true; r.pathid = jstar; r.ob = ostar;
    r.dis_lat = __shfl_sync(DP_FULL, bestd, (nchunk == 1) ? (ostar & 31) : ostar);
    double sum = 0.0;
    for (int t0 = 0; t0 < jstar; t0 += DP_SCR) {
        const int tn = min(DP_SCR, jstar - t0);
        for (int j = lane; j < tn; j += 32) {
            const double2 a = dp_path_point(m, sm, s, t0 + j), b = dp_path_point(m, sm, s, t0 + j + 1);
            sm.scr[j] = sqrt(dp_sq2(b.x - a.x, b.y - a.y));
        }
        __syncwarp();
        sum = dp_seq_sum(sm, tn, sum);
        __syncwarp();
    }
    r.dis_lng = sum;
    return r;
}

// spec_sincos_deg / spec_atan: fixed polynomials shared (as text, not as code) with the oracle.
__device__ __forceinline__ void dp_sincos_deg(double a, double* c, double* s) {
    const double k = rint(a / 90.0);
    const double r = fma(-90.0, k, a);
    const double x = r * (3.14159265358979323846 / 180.0);
    const double x2 = x * x;
    double ps = -1.0 / 355687428096000.0;
    ps = fma(x2, ps, 1.0 / 1307674368000.0);
    ps = fma(x2, ps, -1.0 / 6227020800.0);
    ps = fma(x2, ps, 1.0 / 39916800.0);
    ps = fma(x2, ps, -1.0 / 362880.0);
    ps = fma(x2, ps, 1.0 / 5040.0);
    ps = fma(x2, ps, -1.0 / 120.0);
    ps = fma(x2, ps, 1.0 / 6.0);
    const double sn = fma(-x * x2, ps, x);
    double pc = 1.0 / 20922789888000.0;
    pc = fma(x2, pc, -1.0 / 87178291200.0);
    pc = fma(x2, pc, 1.0 / 479001600.0);
    pc = fma(x2, pc, -1.0 / 3628800.0);
    pc = fma(x2, pc, 1.0 / 40320.0);
    pc = fma(x2, pc, -1.0 / 720.0);
    pc = fma(x2, pc, 1.0 / 24.0);
    pc = fma(x2, pc, -0.5);
    const double cs = fma(x2, pc, 1.0);
    const long long q = (long long)k;
    const int mq = (int)(((q % 4) + 4) % 4);
    if (mq == 0) { *c = cs; *s = sn; }
    else if (mq == 1) { *c = -sn; *s = cs; }
    else if (mq == 2) { *c = -cs; *s = -sn; }
    else { *c = sn; *s = -cs; }
}

__device__ __forceinline__ double dp_atan(double z) {
    const double PI_2 = 1.57079632679489661923, PI_4 = 0.78539816339744830962;
    const bool neg = z < 0;
    double a = neg ? -z : z;
    const bool inv = a > 1.0;
    if (inv) a = 1.0 / a;
    const bool shift = a > 0.41421356237309503;
    const double w = shift ? (a - 1.0) / (a + 1.0) : a;
    const double w2 = w * w;
    double p = 1.0 / 47.0;
#pragma unroll
    for (int n = 22; n >= 0; --n) {
        const double c = 1.0 / (double)(2 * n + 1);
        p = fma(-w2, p, c);
    }
    double r = w * p;
    if (shift) r = PI_4 + r;
    if (inv) r = PI_2 - r;
    return neg ? -r : r;
}

// CalcGlobalDir / GetRoadAngle (Planning.cpp:719-750)
__device__ __forceinline__ double dp_heading(double ax, double ay, double bx, double by, double eps, double pi) {
    double angle;
    if (fabs(bx - ax) < eps && fabs(by - ay) < eps) angle = 0;
    else if (fabs(bx - ax) < eps) angle = (by > ay) ? pi / 2 : 3 * pi / 2;
    else {
        angle = dp_atan((by - ay) / (bx - ax));
        if (bx < ax) angle = angle + pi;
        else if ((bx > ax) && (by < ay)) angle = angle + 2 * pi;
    }
    return angle * 180 / pi;
}

// GetLatDis (Planning.cpp:686-709), LEFT positive
__device__ __forceinline__ double dp_lat_dis(double cx, double cy, double px, double py, double nx, double ny, double eps) {
    double l;
    if (fabs(px - nx) > eps) {
        const double k = (py - ny) / (px - nx);
        l = fabs((cy - py) - k * (cx - px)) / sqrt(1 + k * k);
    } else l = fabs(px - cx);
    if (l < eps) return 0.0;
    const double c = (nx - px) * (cy - py) - (ny - py) * (cx - px);
    return l * (c > 0 ? 1 : -1);
}

// GetAngleErr (Planning.cpp:760-786)
__device__ __forceinline__ double dp_angle_err(double d1, double d2) {
    double e = d2 - d1;
    if (d1 < 180) e = (d2 - d1 <= 180) ? d2 - d1 : d2 - d1 - 360;
    else if (d1 >= 180) e = (d2 - d1 > -180) ? d2 - d1 : d2 - d1 + 360;
    return e;
}

// CShare::BezierPlanning (Planning.cpp:606,863) into sm.plan, one warp
__device__ __forceinline__ void dp_bezier_to_plan(WarpSmem& sm, double x0, double y0, double dir0, double x3, double y3,
                                                  double dir3, int lane) {
    const double ex = x3 - x0, ey = y3 - y0;
    const double L = sqrt(dp_sq2(ex, ey)) / 3.0;
    double c0, s0, c3, s3;
    dp_sincos_deg(dir0, &c0, &s0);
    dp_sincos_deg(dir3, &c3, &s3);
    const double x1 = fma(L, c0, x0), y1 = fma(L, s0, y0);
    const double x2 = fma(-L, c3, x3), y2 = fma(-L, s3, y3);
    for (int i = lane; i < DP_PATH_POINTS; i += 32) {
        const double t = (double)i / (double)(DP_PATH_POINTS - 1);
        const double u = 1.0 - t;
        const double b0 = u * u * u;
        const double b1 = 3.0 * (u * u) * t;
        const double b2 = 3.0 * u * (t * t);
        const double b3 = t * t * t;
        sm.plan[i] = make_double2(fma(b3, x3, fma(b2, x2, fma(b1, x1, b0 * x0))),
                                  fma(b3, y3, fma(b2, y2, fma(b1, y1, b0 * y0))));
    }
    __syncwarp();
}

// CShare::MeanPoints (Planning.cpp:872): resample the first n_in points of `s` into sm.plan.
// n_in <= DP_SCR.  cum[] lives in sm.scr.
__device__ __forceinline__ void dp_mean_points_to_plan(const DevMap& m, WarpSmem& sm, const PathSrc& s, int n_in, int lane) {
    if (n_in <= 0) {
        for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = make_double2(0.0, 0.0);
        __syncwarp();
        return;
    }
    if (n_in == 1) {
        const double2 p = dp_path_point(m, sm, s, 0);
        for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = p;
        __syncwarp();
        return;
    }
    // segment lengths in parallel into tile[].x (n_in - 1 <= 255 terms would not fit the tile: use scr in two passes)
    for (int j = lane; j < n_in - 1; j += 32) {
        const double2 a = dp_path_point(m, sm, s, j), b = dp_path_point(m, sm, s, j + 1);
        sm.scr[j + 1] = sqrt(dp_sq2(b.x - a.x, b.y - a.y));
    }
    __syncwarp();
    if (lane == 0) {                                         // in-place sequential prefix: cum[i+1] = cum[i] + len[i]
        double acc = 0.0;
        sm.scr[0] = 0.0;
        for (int j = 1; j < n_in; ++j) { acc += sm.scr[j]; sm.scr[j] = acc; }
    }
    __syncwarp();
    const double step = sm.scr[n_in - 1] / (double)(DP_PATH_POINTS - 1);
    for (int kx = lane; kx < DP_PATH_POINTS; kx += 32) {
        const double sv = (double)kx * step;
        int lo_i = 0, hi_i = n_in - 2;                       // largest i <= n_in-2 with cum[i] <= sv
        while (lo_i < hi_i) {
            const int mid = (lo_i + hi_i + 1) >> 1;
            if (sm.scr[mid] <= sv) lo_i = mid; else hi_i = mid - 1;
        }
        const int i = lo_i;
        const double seg = sm.scr[i + 1] - sm.scr[i];
        const double t = seg > 0 ? (sv - sm.scr[i]) / seg : 0.0;
        const double2 a = dp_path_point(m, sm, s, i), b = dp_path_point(m, sm, s, i + 1);
        double2 q = make_double2(fma(t, b.x - a.x, a.x), fma(t, b.y - a.y, a.y));
        if (kx == DP_PATH_POINTS - 1) q = dp_path_point(m, sm, s, n_in - 1);
        sm.plan[kx] = q;
    }
    __syncwarp();
}
