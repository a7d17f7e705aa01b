// csrc/dp_fused.cuh -- fused multi-trajectory passes of the Decision half.
//
// The four lane-region trajectories of AroundObstacle (F, R and one neighbour pair, Decision.cpp:811-842)
// are independent, and so are the lateral-offset candidates of the avoid sweep (Decision.cpp:940-974).
// Scoring them one after the other makes a scene a chain of ~8.5 k-cycle searches; here they share ONE
// staging round trip, ONE scan and ONE reduction phase:
//   * dp_region_pass: the <= 320 points of the 4 paths are staged back to back in shared memory by a single
//     coalesced gather; each lane scans one third of that array for its obstacle, keeping one running
//     argmin per path; per-path results are combined across the chunk lanes with shuffles;
//   * dp_sweep_pass: F and its unit normals are staged once; lanes are (candidate, obstacle) pairs, the
//     candidate point p + d*n is rolled out in registers inside the scan (CreateNewPath fused), never stored.
// Results are bit-identical to scoring each trajectory alone with dp_search (tests/test_gpu_parity.py).
#pragma once
#include "dp_device.cuh"

// 4 independent running minima over q[a..b): lexicographic (d2, j) minimum = sequential strict-'<' argmin
__device__ __forceinline__ void dp_scan(const double2* __restrict__ q, int a, int b, double mx, double my, double& bd, int& bj) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double b0 = INF, b1 = INF, b2 = INF, b3 = INF;
    int i0 = a, i1 = a, i2 = a, i3 = a;
    int j = a;
    const double2* p = q + a;
    for (; j + 4 <= b; j += 4, p += 4) {
        const double2 p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
        const double x0 = mx - p0.x, y0 = my - p0.y, x1 = mx - p1.x, y1 = my - p1.y;
        const double x2 = mx - p2.x, y2 = my - p2.y, x3 = mx - p3.x, y3 = my - p3.y;
        const double d0 = fma(x0, x0, y0 * y0), d1 = fma(x1, x1, y1 * y1);
        const double d2 = fma(x2, x2, y2 * y2), d3 = fma(x3, x3, y3 * y3);
        if (d0 < b0) { b0 = d0; i0 = j; }
        if (d1 < b1) { b1 = d1; i1 = j + 1; }
        if (d2 < b2) { b2 = d2; i2 = j + 2; }
        if (d3 < b3) { b3 = d3; i3 = j + 3; }
    }
    for (; j < b; ++j, ++p) {
        const double2 p0 = p[0];
        const double x0 = mx - p0.x, y0 = my - p0.y;
        const double d0 = fma(x0, x0, y0 * y0);
        if (d0 < b0) { b0 = d0; i0 = j; }
    }
    if (b1 < b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
    if (b3 < b2 || (b3 == b2 && i3 < i2)) { b2 = b3; i2 = i3; }
    if (b2 < b0 || (b2 == b0 && i2 < i0)) { b0 = b2; i0 = i2; }
    bd = b0; bj = i0;
}

// gate + signed lateral offset + corridor test of one obstacle against segment (pk, pk1) of a P-point path whose
// nearest point is bj: returns the packed selection key, d in *dout
__device__ __forceinline__ unsigned dp_owner_key(double2 pk, double2 pk1, int bj, int P, int o, double mx, double my, double lo,
                                                 double hi, double* dout) {
    const double sx = pk1.x - pk.x, sy = pk1.y - pk.y;
    bool pass = true;
    if (bj == 0) pass = fma(mx - pk.x, sx, (my - pk.y) * sy) >= 0.0;
    else if (bj == P - 1) pass = fma(mx - pk1.x, sx, (my - pk1.y) * sy) <= 0.0;
    *dout = 0.0;
    if (!pass) return 0xffffffffu;
    const double cross = fma(mx - pk.x, sy, -((my - pk.y) * sx));
    const double len2 = dp_sq2(sx, sy);
    {   // FP32 pre-reject of obstacles far outside the corridor (|d| = |cross| / len): saves the FP64 sqrt and division
        const float cf = (float)cross, lf = (float)len2;
        const float wm = (float)fmax(fabs(lo), fabs(hi)) * 1.001f + 0.01f;
        if (cf * cf > wm * wm * lf * 1.001f) return 0xffffffffu;
    }
    const double len = sqrt(len2);
    double d = 0.0;
    if (len > 0) d = cross / len;
    *dout = d;
    return (d >= lo && d <= hi) ? (((unsigned)bj << 16) | (unsigned)o) : 0xffffffffu;
}

// ---- the four lane-region trajectories share one staging round trip -------------------------------------------
// path r: map points base[r] + stride[r]*j, j < P[r], lateral offset d[r] (virtual neighbour lane).  All (<= 320)
// points are gathered back to back into sm.q by ONE coalesced pass (CreateNewPath fused for virtual lanes), so the
// four SearchObstacle evaluations that follow read shared memory only.
__device__ __forceinline__ void dp_region_stage(const DevMap& m, WarpSmem& sm, const int (&base)[4], const int (&stride)[4],
                                                const int (&P)[4], const double (&d)[4], int lane) {
    double2* Q = sm.q;
    __syncwarp();
    int S = 0;
    // paths 0/2 are forward slices of <= 120 points (4 rounds of 32 lanes), paths 1/3 rear slices of <= 40 (2 rounds): twelve
    // independent predicated gathers with compile-time path parameters (no per-element path lookup)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int Pr = P[r], bs = base[r], st = stride[r];
        const double dr = d[r];
#pragma unroll
        for (int k = 0; k < ((r & 1) ? 2 : 4); ++k) {
            const int j = lane + 32 * k;
            if (j < Pr) {
                double2 pt = m.xy[bs + st * j];
                if (dr != 0.0 && Pr >= 2) {
                    const int jj = min(j, Pr - 2);
                    double2 n = m.nrm[(st > 0) ? bs + jj : bs - jj - 1];
                    if (st < 0) { n.x = -n.x; n.y = -n.y; }
                    pt.x = fma(dr, n.x, pt.x); pt.y = fma(dr, n.y, pt.y);
                }
                Q[S + j] = pt;
            }
        }
        S += Pr;
    }
    __syncwarp();
}

// ---- avoid sweep: several lateral-offset candidates of F per pass ---------------------------------------------------
// stage F (P <= 120 points from map index base) at sm.q[0..P) and the unit normal used for point j at sm.q[128 + j]
__device__ __forceinline__ void dp_sweep_stage(const DevMap& m, WarpSmem& sm, int base, int P, int lane) {
    __syncwarp();
    for (int j = lane; j < P; j += 32) {
        sm.q[j] = m.xy[base + j];
        sm.q[128 + j] = (P >= 2) ? m.nrm[base + min(j, P - 2)] : make_double2(0.0, 0.0);
    }
    __syncwarp();
}
__device__ __forceinline__ double2 dp_sweep_point(const WarpSmem& sm, int j, double dc) {
    const double2 p = sm.q[j], n = sm.q[128 + j];
    return make_double2(fma(dc, n.x, p.x), fma(dc, n.y, p.y));
}
// Candidate 0 of either side is F itself (offsets -0.3*0 and 0.3*0, Decision.cpp:942,961) under the window of the F region
// search: its result is the F region's, already known, and -- the sweep only runs when that gap is < 15 -- never feasible.
// So only the 2(K-1) shifted candidates are scored; they are numbered u = 0 .. 2(K-1)-1 in reference order
// (L1..L(K-1), R1..R(K-1)); dp_sweep_g maps u to the reference's candidate index g (L0..L(K-1), R0..R(K-1)).
__device__ __forceinline__ int dp_sweep_g(int u, int K) { return (u >= K - 1) ? u + 2 : u + 1; }
// lateral offset of candidate g: -0.3*i on the left (g = i < K, Decision.cpp:942), 0.3*i on the right (g = K + i, :961)
__device__ __forceinline__ double dp_sweep_offset(int g, int K) { return (g < K) ? -0.3 * g : 0.3 * (g - K); }
// Scores shifted candidates u0 .. u0+cnt-1 (cnt * N <= 32) against the scene's obstacles (this lane's obstacle =
// (mx, my), N < 32).  sink(g, result, before_first) receives the result of candidate g;
// with need_all == false the arclength of a candidate is only resolved as far as the `> clear` decision needs and
// candidates after the first feasible one are skipped.  Returns the index (0..cnt-1) of the first feasible candidate or -1.
template <class Sink>
// hmax / hmin / dn: pruning bounds of F's lane (longest / shortest segment, largest change of normal); each lane derives the
// bounds of ITS shifted candidate from them (dg_bounds) and runs the exactly pruned scan of dp_group.cuh.
// relmask: the obstacles that can reach the corridor of ANY candidate (bit o), N = their number; lane (ci, r) scores candidate
// u0 + ci against the r-th of them, whose position the caller has put into (mx, my) and whose index into oid -- an obstacle
// left out can only produce the empty key, so the result is the one over all obstacles.
__device__ __forceinline__ int dp_sweep_pass(WarpSmem& sm, int P, int u0, int cnt, int K, double mx, double my, int N, const LaneMap lm,
                                             double lo, double hi, double clear, bool need_all, int lane, Sink sink,
                                             const float hmax, const float hmin, const float dn, const int oid, const unsigned relmask) {
    const int ci_me = lane / N, o = oid;                    // N <= 16 here
    const bool active = ci_me < cnt;
    const int g_me = dp_sweep_g(u0 + ci_me, K);
    const double dc = dp_sweep_offset(g_me, K);
    (void)lm;
    unsigned key = 0xffffffffu;
    double dlat = 0.0;
    if (active && P >= 2) {
        // MY candidate against MY obstacle; the candidate point p + d n is rolled out in registers, never stored
        const float ad = (float)fabs(dc) * 1.0001f;
        const float hb = hmax + ad * dn + 1e-4f;
        float delta = dn;
        if (ad > 0.f) delta = (hmin > 1e-6f) ? dn * (1.0f + 4.0f * ad / hmin) * 1.0001f + 1e-6f : 2.0f;
        const float dmax = (hmax < 0.f) ? dg_inff() : dg_dmax(lo, hi, hb, delta, hmin - ad * dn);   // (hmax < 0: pruning switched off)
        const unsigned long long mk = (hmax < 0.f) ? ~0ull >> (64 - (P + 7) / 8) : dg_coarse_f([&](int j) { return dp_sweep_point(sm, j, dc); }, P, mx, my, hb, dmax);
        DgArg a; a.bd = __longlong_as_double(0x7ff0000000000000LL); a.bj = 0;
        if (mk) a = dg_refine_f([&](int j) { return dp_sweep_point(sm, j, dc); }, P, mx, my, mk);
        const int i0 = a.bj;
        const bool reach = mk && dg_within_reach(a.bd, dmax);
        const int bj = i0, k = (bj == P - 1) ? P - 2 : bj;
        if (reach) key = dp_owner_key(dp_sweep_point(sm, k, dc), dp_sweep_point(sm, k + 1, dc), bj, P, o, mx, my, lo, hi, &dlat);
    }
    // per-candidate selection: min of the packed key over the candidate's N lanes (xor butterfly restricted by compare)
    // done with full-warp shuffles so that every lane takes part: lanes of other candidates are ignored by index
    int first = -1;
#pragma unroll 1
    for (int ci = 0; ci < cnt; ++ci) {
        if (first >= 0 && !need_all) break;
        const unsigned mine = (ci_me == ci) ? key : 0xffffffffu;
        const unsigned gmin = __reduce_min_sync(DP_FULL, mine);
        SearchRes r;
        r.found = false; r.dis_lat = DP_NOT_FOUND; r.dis_lng = DP_NOT_FOUND; r.ob = -1; r.pathid = 0;
        if (P >= 2 && gmin != 0xffffffffu) {
            const int jstar = (int)(gmin >> 16), ostar = (int)(gmin & 0xffffu);
            r.found = true; r.pathid = jstar; r.ob = ostar;
            r.dis_lat = __shfl_sync(DP_FULL, dlat, ci * N + __popc(relmask & ((1u << ostar) - 1u)));
            const int gc = dp_sweep_g(u0 + ci, K);
            const double dcc = dp_sweep_offset(gc, K);
            // Only the `dis_lng > clear` decision is consumed unless a trace is kept.  The jstar segments of the shifted line
            // are each between tl = hmin - |d| dn and th = hmax + |d| dn long (hmin / hmax carry a 1e-4 relative margin, far above
            // the rounding of jstar <= 120 sequential additions), so jstar tl > clear or jstar th <= clear settles it without a sum;
            // on a 0.5 m map with clear = 25 that leaves jstar = 50 for the exact evaluation
            int decided = 0;                                // 1: passes, 2: fails
            if (!need_all && hmax >= 0.f) {
                const float adc = (float)fabs(dcc) * 1.0001f;
                const float tl = hmin - adc * dn, th = hmax + adc * dn;
                if (tl > 0.f && (double)jstar * (double)tl > clear) decided = 1;
                else if ((double)jstar * (double)th <= clear) decided = 2;
            }
            if (decided) r.dis_lng = (decided == 1) ? DP_NOT_FOUND : 0.0;   // (the value itself is not consumed)
            else {
                __syncwarp();
                for (int j = lane; j < jstar; j += 32) {
                    const double2 a = dp_sweep_point(sm, j, dcc), b = dp_sweep_point(sm, j + 1, dcc);
                    sm.scr[j] = sqrt(dp_sq2(b.x - a.x, b.y - a.y));
                }
                dp_pad_scr(sm, jstar, lane);
                __syncwarp();
                if (need_all) r.dis_lng = dp_seq_sum(sm, jstar, 0.0);
                else {                                      // monotone partial sums, stop at the first one beyond the threshold
                    const SeqHit hq = dp_seq_first(sm, jstar, 0.0, 0.0, clear);
                    r.dis_lng = (hq.k >= 0) ? DP_NOT_FOUND : hq.acc;   // (value beyond the threshold is not consumed)
                }
            }
        }
        sink(dp_sweep_g(u0 + ci, K), r, first < 0);          // (candidate, result, scored by the reference too?)
        if (first < 0 && r.dis_lng > clear) first = ci;
    }
    return first;
}
