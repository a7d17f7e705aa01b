// oracle/cshare_spec.cpp -- TEST INFRASTRUCTURE (CPU oracle); see cshare_spec.h for the contract.
// Compile with -ffp-contract=off: every fused multiply-add below is written as fma().
#include "cshare_spec.h"
#include <cmath>

namespace spec {

static inline double sq2(double dx, double dy) { return std::fma(dx, dx, dy * dy); }

double calc_distance(P2 a, P2 b) {
    double dx = a.x - b.x, dy = a.y - b.y;
    return std::sqrt(dx * dx + dy * dy);
}

double spec_atan(double z) {
    const double PI_2 = 1.57079632679489661923, PI_4 = 0.78539816339744830962;
    bool neg = z < 0;
    double a = neg ? -z : z;
    bool inv = a > 1.0;
    if (inv) a = 1.0 / a;
    bool shift = a > 0.41421356237309503;
    double w = shift ? (a - 1.0) / (a + 1.0) : a;
    double w2 = w * w;
    // atan(w) = w * sum_{n=0}^{23} (-1)^n w^(2n) / (2n+1), Horner from the highest term
    double p = 1.0 / 47.0;
    for (int n = 22; n >= 0; --n) {
        double c = 1.0 / (double)(2 * n + 1);
        p = std::fma(-w2, p, c);
    }
    double r = w * p;
    if (shift) r = PI_4 + r;
    if (inv) r = PI_2 - r;
    return neg ? -r : r;
}

double calc_global_dir(P2 a, P2 b, double eps, double pi) {
    double angle;
    double dx = b.x - a.x, dy = b.y - a.y;
    if (std::fabs(dx) < eps && std::fabs(dy) < eps) {
        angle = 0;
    } else if (std::fabs(dx) < eps) {
        angle = (b.y > a.y) ? pi / 2 : 3 * pi / 2;
    } else {
        angle = spec_atan(dy / dx);
        if (b.x < a.x) angle = angle + pi;
        else if ((b.x > a.x) && (b.y < a.y)) angle = angle + 2 * pi;
    }
    return angle * 180 / pi;
}

int nearest_id(P2 q, const P2* pts, int n) {
    double best = 9999;
    int id = 0;
    for (int i = 0; i < n; ++i) {
        double d = calc_distance(q, pts[i]);
        if (d < best) { best = d; id = i; }
    }
    return id;
}

double lat_dis(P2 cur, P2 pt, P2 nx, double eps) {
    double l = 0;
    if (std::fabs(pt.x - nx.x) > eps) {
        double k = (pt.y - nx.y) / (pt.x - nx.x);
        l = std::fabs((cur.y - pt.y) - k * (cur.x - pt.x)) / std::sqrt(1 + k * k);
    } else {
        l = std::fabs(pt.x - cur.x);
    }
    if (l < eps) return 0;
    double c = (nx.x - pt.x) * (cur.y - pt.y) - (nx.y - pt.y) * (cur.x - pt.x);
    return l * (c > 0 ? 1 : -1);
}

void spec_sincos_deg(double a, double* c, double* s) {
    double k = std::rint(a / 90.0);
    double r = std::fma(-90.0, k, a);
    double x = r * (3.14159265358979323846 / 180.0);
    double x2 = x * x;
    // sin x = x * (1 - x2/3! + x2^2/5! - ... - x2^8/17!)
    double ps = -1.0 / 355687428096000.0;                 // -1/17!
    ps = std::fma(x2, ps, 1.0 / 1307674368000.0);         // +1/15!
    ps = std::fma(x2, ps, -1.0 / 6227020800.0);           // -1/13!
    ps = std::fma(x2, ps, 1.0 / 39916800.0);              // +1/11!
    ps = std::fma(x2, ps, -1.0 / 362880.0);               // -1/9!
    ps = std::fma(x2, ps, 1.0 / 5040.0);                  // +1/7!
    ps = std::fma(x2, ps, -1.0 / 120.0);                  // -1/5!
    ps = std::fma(x2, ps, 1.0 / 6.0);                     // +1/3!  (sign folded below)
    double sn = std::fma(-x * x2, ps, x);
    // cos x = 1 - x2/2! + x2^2/4! - ... + x2^8/16!
    double pc = 1.0 / 20922789888000.0;                   // +1/16!
    pc = std::fma(x2, pc, -1.0 / 87178291200.0);          // -1/14!
    pc = std::fma(x2, pc, 1.0 / 479001600.0);             // +1/12!
    pc = std::fma(x2, pc, -1.0 / 3628800.0);              // -1/10!
    pc = std::fma(x2, pc, 1.0 / 40320.0);                 // +1/8!
    pc = std::fma(x2, pc, -1.0 / 720.0);                  // -1/6!
    pc = std::fma(x2, pc, 1.0 / 24.0);                    // +1/4!
    pc = std::fma(x2, pc, -0.5);                          // -1/2!
    double cs = std::fma(x2, pc, 1.0);
    long long q = (long long)k;
    int m = (int)(((q % 4) + 4) % 4);
    switch (m) {
        case 0: *c = cs;  *s = sn;  break;
        case 1: *c = -sn; *s = cs;  break;
        case 2: *c = -cs; *s = -sn; break;
        default: *c = sn; *s = -cs; break;
    }
}

SearchResult search_obstacle(const P2* p, int P, const P2* obs, int N, double lat_min, double lat_max) {
    SearchResult r{false, NOT_FOUND, NOT_FOUND, -1, 0};
    if (P < 2) return r;
    int best_j = P;  // smaller is better
    for (int o = 0; o < N; ++o) {
        double ox = obs[o].x, oy = obs[o].y;
        double bd = sq2(ox - p[0].x, oy - p[0].y);
        int bj = 0;
        for (int j = 1; j < P; ++j) {
            double d = sq2(ox - p[j].x, oy - p[j].y);
            if (d < bd) { bd = d; bj = j; }
        }
        int k = (bj == P - 1) ? P - 2 : bj;
        double sx = p[k + 1].x - p[k].x, sy = p[k + 1].y - p[k].y;
        if (bj == 0) {
            double dot = std::fma(ox - p[0].x, sx, (oy - p[0].y) * sy);
            if (!(dot >= 0)) continue;
        } else if (bj == P - 1) {
            double dot = std::fma(ox - p[P - 1].x, sx, (oy - p[P - 1].y) * sy);
            if (!(dot <= 0)) continue;
        }
        double len = std::sqrt(sq2(sx, sy));
        double d = 0;
        if (len > 0) d = std::fma(ox - p[k].x, sy, -((oy - p[k].y) * sx)) / len;
        if (!(d >= lat_min && d <= lat_max)) continue;
        if (bj < best_j) {
            best_j = bj;
            r.found = true;
            r.dis_lat = d;
            r.ob_index = o;
            r.pathid = bj;
        }
    }
    if (r.found) {
        double s = 0;
        for (int i = 0; i < r.pathid; ++i) s += std::sqrt(sq2(p[i + 1].x - p[i].x, p[i + 1].y - p[i].y));
        r.dis_lng = s;
    }
    return r;
}

SearchResult search_obstacle_tracks(const P2* p, int P, const P2* obs, const P2* dv, int N, double lat_min, double lat_max) {
    SearchResult r{false, NOT_FOUND, NOT_FOUND, -1, 0};
    if (P < 2) return r;
    int best_j = P;
    for (int o = 0; o < N; ++o) {
        const double vx = dv ? dv[o].x : 0.0, vy = dv ? dv[o].y : 0.0;
        double bd = 0; int bj = 0;
        for (int j = 0; j < P; ++j) {
            double ox = std::fma((double)j, vx, obs[o].x), oy = std::fma((double)j, vy, obs[o].y);
            double d = sq2(ox - p[j].x, oy - p[j].y);
            if (j == 0 || d < bd) { bd = d; bj = j; }
        }
        double ox = std::fma((double)bj, vx, obs[o].x), oy = std::fma((double)bj, vy, obs[o].y);
        int k = (bj == P - 1) ? P - 2 : bj;
        double sx = p[k + 1].x - p[k].x, sy = p[k + 1].y - p[k].y;
        if (bj == 0) {
            double dot = std::fma(ox - p[0].x, sx, (oy - p[0].y) * sy);
            if (!(dot >= 0)) continue;
        } else if (bj == P - 1) {
            double dot = std::fma(ox - p[P - 1].x, sx, (oy - p[P - 1].y) * sy);
            if (!(dot <= 0)) continue;
        }
        double len = std::sqrt(sq2(sx, sy));
        double d = 0;
        if (len > 0) d = std::fma(ox - p[k].x, sy, -((oy - p[k].y) * sx)) / len;
        if (!(d >= lat_min && d <= lat_max)) continue;
        if (bj < best_j) { best_j = bj; r.found = true; r.dis_lat = d; r.ob_index = o; r.pathid = bj; }
    }
    if (r.found) {
        double s = 0;
        for (int i = 0; i < r.pathid; ++i) s += std::sqrt(sq2(p[i + 1].x - p[i].x, p[i + 1].y - p[i].y));
        r.dis_lng = s;
    }
    return r;
}

void rollout_ctr(double x0, double y0, double vx, double vy, double dtheta_deg, int T, double* out_x, double* out_y, int stride) {
    double c, s;
    spec_sincos_deg(dtheta_deg, &c, &s);
    double x = x0, y = y0;
    for (int j = 0; j < T; ++j) {
        out_x[(size_t)j * stride] = x; out_y[(size_t)j * stride] = y;
        x = x + vx; y = y + vy;
        const double nvx = std::fma(c, vx, -(s * vy)), nvy = std::fma(s, vx, c * vy);
        vx = nvx; vy = nvy;
    }
}

SearchResult search_obstacle_tile(const P2* p, int P, const double* tx, const double* ty, int T, int N, double lat_min, double lat_max) {
    SearchResult r{false, NOT_FOUND, NOT_FOUND, -1, 0};
    if (P < 2 || T < 1) return r;
    int best_j = P;
    for (int o = 0; o < N; ++o) {
        double bd = 0; int bj = 0;
        for (int j = 0; j < P; ++j) {
            const int jt = j < T ? j : T - 1;
            const double ox = tx[(size_t)jt * N + o], oy = ty[(size_t)jt * N + o];
            double d = sq2(ox - p[j].x, oy - p[j].y);
            if (j == 0 || d < bd) { bd = d; bj = j; }
        }
        const int jt = bj < T ? bj : T - 1;
        const double ox = tx[(size_t)jt * N + o], oy = ty[(size_t)jt * N + o];
        int k = (bj == P - 1) ? P - 2 : bj;
        double sx = p[k + 1].x - p[k].x, sy = p[k + 1].y - p[k].y;
        if (bj == 0) {
            double dot = std::fma(ox - p[0].x, sx, (oy - p[0].y) * sy);
            if (!(dot >= 0)) continue;
        } else if (bj == P - 1) {
            double dot = std::fma(ox - p[P - 1].x, sx, (oy - p[P - 1].y) * sy);
            if (!(dot <= 0)) continue;
        }
        double len = std::sqrt(sq2(sx, sy));
        double d = 0;
        if (len > 0) d = std::fma(ox - p[k].x, sy, -((oy - p[k].y) * sx)) / len;
        if (!(d >= lat_min && d <= lat_max)) continue;
        if (bj < best_j) { best_j = bj; r.found = true; r.dis_lat = d; r.ob_index = o; r.pathid = bj; }
    }
    if (r.found) {
        double s = 0;
        for (int i = 0; i < r.pathid; ++i) s += std::sqrt(sq2(p[i + 1].x - p[i].x, p[i + 1].y - p[i].y));
        r.dis_lng = s;
    }
    return r;
}

void create_new_path(const P2* p, int P, double d, P2* out) {
    if (P < 2) {
        for (int j = 0; j < P; ++j) out[j] = p[j];
        return;
    }
    for (int j = 0; j < P; ++j) {
        int k = (j == P - 1) ? P - 2 : j;
        double sx = p[k + 1].x - p[k].x, sy = p[k + 1].y - p[k].y;
        double len = std::sqrt(sq2(sx, sy));
        double nx = 0, ny = 0;
        if (len > 0) { nx = sy / len; ny = -sx / len; }
        out[j].x = std::fma(d, nx, p[j].x);
        out[j].y = std::fma(d, ny, p[j].y);
    }
}

void bezier_planning(P3 a, P3 b, P2* out, int n) {
    double ex = b.x - a.x, ey = b.y - a.y;
    double D = std::sqrt(sq2(ex, ey));
    double L = D / 3.0;
    double c0, s0, c3, s3;
    spec_sincos_deg(a.dir, &c0, &s0);
    spec_sincos_deg(b.dir, &c3, &s3);
    double x0 = a.x, y0 = a.y, x3 = b.x, y3 = b.y;
    double x1 = std::fma(L, c0, x0), y1 = std::fma(L, s0, y0);
    double x2 = std::fma(-L, c3, x3), y2 = std::fma(-L, s3, y3);
    for (int i = 0; i < n; ++i) {
        double t = (n > 1) ? (double)i / (double)(n - 1) : 0.0;
        double u = 1.0 - t;
        double b0 = u * u * u;
        double b1 = 3.0 * (u * u) * t;
        double b2 = 3.0 * u * (t * t);
        double b3 = t * t * t;
        out[i].x = std::fma(b3, x3, std::fma(b2, x2, std::fma(b1, x1, b0 * x0)));
        out[i].y = std::fma(b3, y3, std::fma(b2, y2, std::fma(b1, y1, b0 * y0)));
    }
}

void mean_points(const P2* in, int n_in, P2* out, int n_out) {
    if (n_in <= 0) {
        for (int k = 0; k < n_out; ++k) out[k] = P2{0, 0};
        return;
    }
    if (n_in == 1) {
        for (int k = 0; k < n_out; ++k) out[k] = in[0];
        return;
    }
    // cum[] is kept on the stack in chunks: n_in is bounded by the caller (<= 4096 here)
    static thread_local double cum[4096];
    if (n_in > 4096) n_in = 4096;
    cum[0] = 0;
    for (int i = 0; i + 1 < n_in; ++i)
        cum[i + 1] = cum[i] + std::sqrt(sq2(in[i + 1].x - in[i].x, in[i + 1].y - in[i].y));
    double step = (n_out > 1) ? cum[n_in - 1] / (double)(n_out - 1) : 0.0;
    int i = 0;
    for (int k = 0; k < n_out; ++k) {
        double s = (double)k * step;
        while (i < n_in - 2 && cum[i + 1] <= s) ++i;
        double seg = cum[i + 1] - cum[i];
        double t = seg > 0 ? (s - cum[i]) / seg : 0.0;
        out[k].x = std::fma(t, in[i + 1].x - in[i].x, in[i].x);
        out[k].y = std::fma(t, in[i + 1].y - in[i].y, in[i].y);
    }
    out[n_out - 1] = in[n_in - 1];
}

void global_to_wgs84(const Datum& d, double x, double y, double* lat, double* lng) {
    *lat = std::fma(y, d.k_lat, d.lat0);
    *lng = std::fma(x, d.k_lng, d.lng0);
}
void wgs84_to_global(const Datum& d, double lat, double lng, double* x, double* y) {
    *y = (lat - d.lat0) / d.k_lat;
    *x = (lng - d.lng0) / d.k_lng;
}

}  // namespace spec
