// csrc/dp_device.cuh -- device-side geometry operators (the CShare seam of the reference,
// SURVEY.md 8a rows A1-A5) for sm_100a.  One WARP cooperates on one path/trajectory:
//   * lanes are (obstacle, point-chunk) work items for the nearest-point search, so that
//     N = 10 obstacles still fill 30 of 32 lanes; N >= 32 runs in groups of 32 obstacles;
//   * a path is a RECIPE (map slice forward/reversed, optional lateral offset, junction
//     concatenation, or a window of the local path): its points are staged per warp in a
//     shared-memory tile by one coalesced gather from the L2-resident AoS map with the offset
//     fused in, then read as 16-byte broadcast loads -- candidate paths never touch HBM;
//   * selection is a packed (path index << 16 | obstacle index) warp min-reduction (redux.sync),
//     i.e. nearest-along-path first, lowest obstacle index on ties;
//   * sums the reference evaluates sequentially (arclength) stay sequential: terms are produced
//     in parallel into shared memory, then added in index order in blocks of 8 (loads and exit
//     tests off the dependency chain), so results are bit-identical to a scalar CPU evaluation.
// Arithmetic: IEEE binary64, compiled with -fmad=false; fma() appears exactly where the
// operator specification says so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dmpp_b200.h"
#include "dp_group.cuh"

#define DP_FULL 0xffffffffu
#define DP_TILE 120            // path points staged per warp per tile (= the reference's 120-point lane slice)
#define DP_SCR 192             // sequential-sum terms per pass (multiple of 8)
#ifndef DP_WARPS_PER_BLOCK
#define DP_WARPS_PER_BLOCK 4                   // operator kernels; the cycle kernel is templated on its CTA size (dp_cycle.cu)
#endif

// the map tables (+ pruning bounds, prefix / run-end tables) are shared with the group kernel
typedef DgMap DevMap;

// 6864 bytes per warp (+ 1 KB the system reserves per CTA): 7 CTAs of 4 warps (28 warps) or 25 one-warp CTAs fit the
// 196 KB shared-memory configuration and leave 60 KB of L1 for the map gathers.
struct __align__(16) WarpSmem {
    union {
        struct {
            double2 plan[DP_PATH_POINTS];      // last_Bpoints (TMA bulk copy) during the planning half, road_points on exit
            double2 tile[DP_TILE];             // one staged path tile (also cum[] of MeanPoints, 240 doubles)
        };
        double2 q[DP_PATH_POINTS + DP_TILE];   // decision half: the 4 lane-region paths (<= 320 points) staged back to back,
    };                                         //                or F + its normals for the fused avoid sweep
    double scr[DP_SCR + 8];                    // sequential-sum terms, zero padded to a multiple of 8
    unsigned long long mbar;                   // mbarrier of the bulk copy
    unsigned long long pad_;
    uint32_t hdr[32];                          // the scene's dp_scene_hdr, fetched by ONE coalesced 128-byte warp load
};

// A path is a recipe, never an array in HBM: up to two runs of map (or caller) points read with a
// stride of +1/-1, an optional lateral offset applied with the precomputed unit normals while the
// tile is staged, or a window of the local path already resident in sm.plan.
struct Src {
    const double2* p0; int stride0; int n0;    // run 0: p0[stride0 * j], j < n0
    const double2* p1; int n1;                 // run 1 (forward): p1[j - n0]
    const double2* nrm0;                       // normals aligned with p0 (only when d != 0, single run)
    double d;                                  // lateral offset, RIGHT positive
    int q_off;                                 // >= 0: the path is already staged at sm.q[q_off ...] (no staging needed)
};
struct LaneMap { int nchunk, o, c; unsigned magic; };   // nearest-search work item of this lane for N < 32; magic: x / nchunk
// x / lm.nchunk for 0 <= x < 8192 without an integer division (nchunk <= 32; checked exhaustively on the host)
__device__ __forceinline__ int dp_div_chunks(int x, const LaneMap& lm) { return (int)(((unsigned)x * lm.magic) >> 20); }
struct SearchRes { bool found; double dis_lat, dis_lng; int ob, pathid; };

__device__ __forceinline__ LaneMap dp_lane_map(int N, int lane) {
    LaneMap lm;
    if (N >= 32 || N <= 0) { lm.nchunk = 1; lm.o = lane; lm.c = 0; }
    else { lm.nchunk = 32 / N; lm.c = lane / N; lm.o = lane - lm.c * N; }
    lm.magic = (1u << 20) / (unsigned)lm.nchunk + 1u;
    return lm;
}
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
__device__ __forceinline__ Src dp_src_run(const double2* p, int stride, int n) {
    Src s;
    s.p0 = p; s.stride0 = stride; s.n0 = n; s.p1 = p; s.n1 = 0; s.nrm0 = nullptr; s.d = 0.0; s.q_off = -1;
    return s;
}
// point j of the path (CreateNewPath fused in: p_j + d * n_j; fwd normal index min(j,P-2),
// reversed: -(nrm[-min(j,P-2)-1]))
__device__ __forceinline__ double2 dp_src_point(const Src& s, int j) {
    if (j >= s.n0) return s.p1[j - s.n0];
    double2 q = s.p0[s.stride0 * j];
    if (s.d != 0.0 && s.n0 >= 2) {
        const int jj = min(j, s.n0 - 2);
        double2 n = (s.stride0 > 0) ? s.nrm0[jj] : s.nrm0[-jj - 1];
        if (s.stride0 < 0) { n.x = -n.x; n.y = -n.y; }
        q.x = fma(s.d, n.x, q.x); q.y = fma(s.d, n.y, q.y);
    }
    return q;
}

// ---- sequential sums over sm.scr[0..n) in index order, 8 terms per block ------------------------
__device__ __forceinline__ void dp_pad_scr(WarpSmem& sm, int n, int lane) {   // zero terms up to the next multiple of 8
    if (lane < 8) sm.scr[n + lane] = 0.0;
}
__device__ __forceinline__ double dp_seq_sum(const WarpSmem& sm, int n, double acc) {
#pragma unroll 1
    for (int j = 0; j < n; j += 8) {
        const double t0 = sm.scr[j], t1 = sm.scr[j + 1], t2 = sm.scr[j + 2], t3 = sm.scr[j + 3];
        const double t4 = sm.scr[j + 4], t5 = sm.scr[j + 5], t6 = sm.scr[j + 6], t7 = sm.scr[j + 7];
        acc += t0; acc += t1; acc += t2; acc += t3; acc += t4; acc += t5; acc += t6; acc += t7;   // x + 0.0 == x for the padding
    }
    return acc;
}
// first index j < n whose running sum s_j = acc + t_0 + ... + t_j satisfies (s_j - sub > thr), else -1 (acc advanced by
// all n terms).  Terms are >= 0, so the test is monotone in j: a block of 8 terms is only searched when its LAST partial
// sum passes the test -- one compare per block instead of eight.
struct SeqHit { int k; double acc; };
__device__ __forceinline__ SeqHit dp_seq_first(const WarpSmem& sm, int n, double acc, double sub, double thr) {
    SeqHit r;
#pragma unroll 1
    for (int j = 0; j < n; j += 8) {
        const double t0 = sm.scr[j], t1 = sm.scr[j + 1], t2 = sm.scr[j + 2], t3 = sm.scr[j + 3];
        const double t4 = sm.scr[j + 4], t5 = sm.scr[j + 5], t6 = sm.scr[j + 6], t7 = sm.scr[j + 7];
        const double s0 = acc + t0, s1 = s0 + t1, s2 = s1 + t2, s3 = s2 + t3, s4 = s3 + t4, s5 = s4 + t5, s6 = s5 + t6, s7 = s6 + t7;
        if ((s7 - sub) > thr) {
            int k = j + 7;
            if ((s6 - sub) > thr) k = j + 6;
            if ((s5 - sub) > thr) k = j + 5;
            if ((s4 - sub) > thr) k = j + 4;
            if ((s3 - sub) > thr) k = j + 3;
            if ((s2 - sub) > thr) k = j + 2;
            if ((s1 - sub) > thr) k = j + 1;
            if ((s0 - sub) > thr) k = j;
            if (k < n) { r.k = k; r.acc = acc; return r; }   // (padding terms are zero: they cannot create a first hit)
        }
        acc = s7;
    }
    r.k = -1; r.acc = acc;
    return r;
}

// ---- CShare::NearestId / the nearest-point loop of GetVhclLocalState (Planning.cpp:640-648), one warp -----------------
// Reference: d_i = sqrt(e_i) with e_i = dx*dx + dy*dy, strict '<' starting from 9999: the LOWEST index among the points
// whose ROUNDED distance is minimal (0x7fffffff here when no distance is below 9999).  sqrt is monotone, so that distance
// is sqrt(min e) and the answer is the lowest index with e == min e -- unless a DIFFERENT e within a few ulps of the
// minimum rounds to the same sqrt.  So the squared distances are compared (one sqrt per query instead of n), and any e
// within 2^-49 of the minimum but not equal to it sends the warp through the reference's own loop.
__device__ __forceinline__ int dp_nearest_plain(const double2* pts, int n, double x, double y, int lane) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL), TOL = 1.0 + 0x1p-49;
    double be = INF;
    int bi = 0x7fffffff;
    bool amb = false;
    for (int i = lane; i < n; i += 32) {
        const double2 q = pts[i];
        const double dx = x - q.x, dy = y - q.y;
        const double e = dx * dx + dy * dy;                 // the argument of the reference's sqrt, same operations
        if (e < be) { amb = amb || (be <= e * TOL); be = e; bi = i; }   // (the displaced lower index may round to the same sqrt)
    }
    double ge = be;
    int gi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oe = __shfl_xor_sync(DP_FULL, ge, o);
        const int oi = __shfl_xor_sync(DP_FULL, gi, o);
        if (oe < ge || (oe == ge && oi < gi)) { ge = oe; gi = oi; }
    }
    amb = amb || (be != ge && be <= ge * TOL);
    if (!__any_sync(DP_FULL, amb)) return (gi != 0x7fffffff && sqrt(ge) < 9999.0) ? gi : 0x7fffffff;
    double md = 9999.0;                                     // near tie: the reference loop, verbatim
    int mi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const double2 q = pts[i];
        const double dd = dp_dist_plain(x, y, q.x, q.y);
        if (dd < md) { md = dd; mi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(DP_FULL, md, o);
        const int oi = __shfl_xor_sync(DP_FULL, mi, o);
        if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
    }
    return mi;
}

// ---- TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) with an mbarrier --------------
__device__ __forceinline__ uint32_t dp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dp_bulk_prefetch(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* mbar, int lane) {
    if (lane == 0) {
        const uint32_t mb = dp_smem_u32(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dp_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(mb) : "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ void dp_bulk_wait(unsigned long long* mbar) {
    const uint32_t mb = dp_smem_u32(mbar);
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(mb) : "memory");
    }
    if (!done) __trap();                                    // never spin forever on a broken copy
}

// ---- CShare::SearchObstacle (Planning.cpp:168; Decision.cpp:370,455,811-842,943,962), one warp ----
// One copy of this code serves every trajectory of the cycle (noinline keeps the kernel inside the
// instruction cache).  (mx, my): this lane's obstacle when N < 32 (kept in registers for the whole
// cycle); ox/oy: the scene's obstacle rows in global memory for the grouped N >= 32 case.
// hb > 0 switches on the exactly pruned scan of dp_group.cuh (hb >= the path's segment lengths, dmax = reach of the corridor):
// a lane samples one point per cell of 8 of ITS chunk and scans only the cells that can hold a point within
// min(best sample, dmax); an obstacle whose best point is farther than dmax cannot be in the corridor and is dropped.
__device__ __forceinline__ SearchRes dp_search(const Src s, double mx, double my, const double* __restrict__ ox,
                                                   const double* __restrict__ oy, int N, const LaneMap lm, double lo, double hi,
                                                   WarpSmem& sm, int lane, const float hb = 0.f, const float dmax = 0.f,
                                                   double* bd_out = nullptr) {
    // bd_out (N < 32 only): lane o < N receives the squared distance from obstacle o to its nearest path point (+inf: none)
    if (bd_out) *bd_out = __longlong_as_double(0x7ff0000000000000LL);
    SearchRes r;
    r.found = false; r.dis_lat = DP_NOT_FOUND; r.dis_lng = DP_NOT_FOUND; r.ob = -1; r.pathid = 0;
    const int P = s.n0 + s.n1;
    if (P < 2 || N <= 0) return r;
    const int nchunk = lm.nchunk;
    const int CS = (nchunk == 1 || P >= 8192) ? ((nchunk == 1) ? P : (P + nchunk - 1) / nchunk) : dp_div_chunks(P + nchunk - 1, lm);
    const int ngroups = (nchunk == 1) ? (N + 31) >> 5 : 1;
    const bool in_plan = s.q_off >= 0;
    const bool one_tile = in_plan || P <= DP_TILE;          // every point stays addressable in shared memory
    const double2* pts = &sm.q[in_plan ? s.q_off : DP_PATH_POINTS];      // shared: LDS in the hot loop (tile = q[200..320))
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    unsigned bestkey = 0xffffffffu;
    double bestd = 0.0;
    for (int g = 0; g < ngroups; ++g) {
        const int o = (nchunk == 1) ? g * 32 + lane : lm.o;
        const bool active = (nchunk == 1) ? (o < N) : (lane < N * nchunk);
        if (nchunk == 1) { mx = active ? __ldcg(ox + o) : 0.0; my = active ? __ldcg(oy + o) : 0.0; }   // (L2: rows may have been staged by the overlapped Decision launch)
        const int jlo = lm.c * CS, jhi = active ? min(P, jlo + CS) : 0;
        // four independent running minima (j mod 4 classes) break the compare/select dependency chain;
        // their lexicographic (d2, j) minimum is the sequential strict-'<' result (lowest index on ties)
        double b0 = INF, b1 = INF, b2 = INF, b3 = INF;
        int i0 = jlo, i1 = jlo, i2 = jlo, i3 = jlo;
        const int tstep = in_plan ? P : DP_TILE;
        const bool prune = hb > 0.f && one_tile;
        if (prune) {
            if (!in_plan && g == 0) {                       // stage the (single) tile: coalesced gather, offset fused
                __syncwarp();
                for (int j = lane; j < P; j += 32) sm.tile[j] = dp_src_point(s, j);
                __syncwarp();
            }
            const int len = jhi - jlo;
            if (len > 0) {
                const double2* q = pts + jlo;
                const unsigned long long mk = dg_coarse_f([&](int j) { return q[j]; }, len, mx, my, hb, dmax);
                if (mk) {
                    const DgArg a = dg_refine_f([&](int j) { return q[j]; }, len, mx, my, mk);
                    b0 = a.bd; i0 = jlo + a.bj;
                }
            }
        }
        for (int t0 = 0; t0 < P && !prune; t0 += tstep) {
            const int tn = min(tstep, P - t0);
            if (!in_plan && (g == 0 || !one_tile)) {        // stage the tile: coalesced gather, offset fused
                __syncwarp();
                for (int j = lane; j < tn; j += 32) sm.tile[j] = dp_src_point(s, t0 + j);
                __syncwarp();
            }
            int j = max(jlo, t0);
            const int e = min(jhi, t0 + tn);
            const double2* q = pts + (j - t0);
#pragma unroll 2
            for (; j + 4 <= e; j += 4, q += 4) {
                const double2 p0 = q[0], p1 = q[1], p2 = q[2], p3 = q[3];
                const double x0 = mx - p0.x, y0 = my - p0.y, x1 = mx - p1.x, y1 = my - p1.y;
                const double x2 = mx - p2.x, y2 = my - p2.y, x3 = mx - p3.x, y3 = my - p3.y;
                const double d0 = fma(x0, x0, y0 * y0), d1 = fma(x1, x1, y1 * y1);
                const double d2 = fma(x2, x2, y2 * y2), d3 = fma(x3, x3, y3 * y3);
                if (d0 < b0) { b0 = d0; i0 = j; }
                if (d1 < b1) { b1 = d1; i1 = j + 1; }
                if (d2 < b2) { b2 = d2; i2 = j + 2; }
                if (d3 < b3) { b3 = d3; i3 = j + 3; }
            }
            for (; j < e; ++j, ++q) {
                const double2 p0 = q[0];
                const double x0 = mx - p0.x, y0 = my - p0.y;
                const double d0 = fma(x0, x0, y0 * y0);
                if (d0 < b0) { b0 = d0; i0 = j; }
            }
        }
        if (b1 < b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
        if (b3 < b2 || (b3 == b2 && i3 < i2)) { b2 = b3; i2 = i3; }
        if (b2 < b0 || (b2 == b0 && i2 < i0)) { b0 = b2; i0 = i2; }
        double bd = b0;
        int bj = i0;
        for (int cc = 1; cc < nchunk; ++cc) {               // chunks are index-ordered: strict '<' keeps the lowest j
            const int src = (lm.o + cc * N) & 31;
            const double od = __shfl_sync(DP_FULL, bd, src);
            const int oj = __shfl_sync(DP_FULL, bj, src);
            if (lane < N && od < bd) { bd = od; bj = oj; }
        }
        if (bd_out && nchunk > 1) *bd_out = bd;
        const bool owner = ((nchunk == 1) ? active : (lane < N)) && (!prune || (bd < INF && dg_within_reach(bd, dmax)));
        if (owner) {
            const int k = (bj == P - 1) ? P - 2 : bj;
            const double2 pk = one_tile ? pts[k] : dp_src_point(s, k), pk1 = one_tile ? pts[k + 1] : dp_src_point(s, k + 1);
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
    for (int t0 = 0; t0 < jstar; t0 += DP_SCR) {            // arclength to the selected point, index order
        const int tn = min(DP_SCR, jstar - t0);
        __syncwarp();
        for (int j = lane; j < tn; j += 32) {
            const double2 a = one_tile ? pts[t0 + j] : dp_src_point(s, t0 + j), b = one_tile ? pts[t0 + j + 1] : dp_src_point(s, t0 + j + 1);
            sm.scr[j] = sqrt(dp_sq2(b.x - a.x, b.y - a.y));
        }
        dp_pad_scr(sm, tn, lane);
        __syncwarp();
        sum = dp_seq_sum(sm, tn, sum);
    }
    r.dis_lng = sum;
    return r;
}

// spec_sincos_deg / spec_atan: fixed polynomials shared (as text, not as code) with the oracle.
static __device__ __noinline__ void dp_sincos_deg(double a, double* c, double* s) {
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

static __device__ __noinline__ double dp_atan(double z) {
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
static __device__ __noinline__ double dp_heading(double ax, double ay, double bx, double by, double eps, double pi) {
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
static __device__ __noinline__ double dp_lat_dis(double cx, double cy, double px, double py, double nx, double ny, double eps) {
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
static __device__ __noinline__ void dp_bezier_to_plan(WarpSmem& sm, double x0, double y0, double dir0, double x3, double y3,
                                                      double dir3, int lane) {
    const double ex = x3 - x0, ey = y3 - y0;
    const double L = sqrt(dp_sq2(ex, ey)) / 3.0;
    double c0, s0, c3, s3;
    dp_sincos_deg(dir0, &c0, &s0);
    dp_sincos_deg(dir3, &c3, &s3);
    const double x1 = fma(L, c0, x0), y1 = fma(L, s0, y0);
    const double x2 = fma(-L, c3, x3), y2 = fma(-L, s3, y3);
    __syncwarp();
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

// CShare::MeanPoints (Planning.cpp:872): resample the first n_in (<= 240) points of `s` into sm.plan.
// cum[] lives in the (idle) tile area.
static __device__ __noinline__ void dp_mean_points_to_plan(WarpSmem& sm, const Src s, int n_in, int lane) {
    double* cum = reinterpret_cast<double*>(sm.tile);
    __syncwarp();
    if (n_in <= 0) {
        for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = make_double2(0.0, 0.0);
        __syncwarp();
        return;
    }
    if (n_in == 1) {
        const double2 p = dp_src_point(s, 0);
        for (int i = lane; i < DP_PATH_POINTS; i += 32) sm.plan[i] = p;
        __syncwarp();
        return;
    }
    for (int j = lane; j < n_in - 1; j += 32) {
        const double2 a = dp_src_point(s, j), b = dp_src_point(s, j + 1);
        cum[j + 1] = sqrt(dp_sq2(b.x - a.x, b.y - a.y));
    }
    __syncwarp();
    if (lane == 0) {                                         // in-place sequential prefix: cum[i+1] = cum[i] + len[i]
        double acc = 0.0;
        cum[0] = 0.0;
        for (int j = 1; j < n_in; ++j) { acc += cum[j]; cum[j] = acc; }
    }
    __syncwarp();
    const double step = cum[n_in - 1] / (double)(DP_PATH_POINTS - 1);
    for (int kx = lane; kx < DP_PATH_POINTS; kx += 32) {
        const double sv = (double)kx * step;
        int lo_i = 0, hi_i = n_in - 2;                       // largest i <= n_in-2 with cum[i] <= sv
        while (lo_i < hi_i) {
            const int mid = (lo_i + hi_i + 1) >> 1;
            if (cum[mid] <= sv) lo_i = mid; else hi_i = mid - 1;
        }
        const int i = lo_i;
        const double seg = cum[i + 1] - cum[i];
        const double t = seg > 0 ? (sv - cum[i]) / seg : 0.0;
        const double2 a = dp_src_point(s, i), b = dp_src_point(s, i + 1);
        double2 q = make_double2(fma(t, b.x - a.x, a.x), fma(t, b.y - a.y, a.y));
        if (kx == DP_PATH_POINTS - 1) q = dp_src_point(s, n_in - 1);
        sm.plan[kx] = q;
    }
    __syncwarp();
}
