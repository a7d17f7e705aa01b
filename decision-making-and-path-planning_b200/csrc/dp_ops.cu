// csrc/dp_ops.cu -- operator-level kernels behind the CShare seam, the
// dense candidate sweep of BASELINE config 3, and the FMA micro-benchmark that provides the
// roofline denominator (MEASURED_PEAKS.json has no FP64/FP32 CUDA-core figure).
#include <cstdlib>
#include <cstring>
#include <cooperative_groups.h>
#include "dp_device.cuh"
#include "dp_fused.cuh"
#include "dp_kernels.h"

namespace {

// CShare::SearchObstacle, one warp per path
__global__ void __launch_bounds__(DP_WARPS_PER_BLOCK * 32)
op_search_kernel(int n_paths, const int32_t* __restrict__ path_off, const double2* __restrict__ pxy, const double* __restrict__ ox,
                 const double* __restrict__ oy, int n_obs, const double* __restrict__ lat_min, const double* __restrict__ lat_max,
                 dp_search_slot* __restrict__ out) {
    __shared__ WarpSmem smem[DP_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pid = blockIdx.x * DP_WARPS_PER_BLOCK + wib;
    if (pid >= n_paths) return;
    const int off = path_off[pid], P = path_off[pid + 1] - off;
    const LaneMap lm = dp_lane_map(n_obs, lane);
    const bool act = (lm.nchunk > 1) && (lane < n_obs * lm.nchunk);
    const SearchRes r = dp_search(dp_src_run(pxy + off, 1, P), act ? ox[lm.o] : 0.0, act ? oy[lm.o] : 0.0, ox, oy, n_obs, lm, lat_min[pid],
                                  lat_max[pid], smem[wib], lane);
    if (lane == 0) {
        dp_search_slot o;
        o.dis_lat = r.dis_lat; o.dis_lng = r.dis_lng; o.ob_index = (int16_t)r.ob; o.pathid = (uint16_t)r.pathid;
        o.evaluated = 1; o.found = r.found; o.pad[0] = o.pad[1] = 0;
        out[pid] = o;
    }
}

// CShare::CreateNewPath, one warp per path
__global__ void op_create_kernel(int n_paths, const int32_t* __restrict__ path_off, const double2* __restrict__ pxy,
                                 const double* __restrict__ offset, double2* __restrict__ out_xy) {
    const int lane = threadIdx.x & 31;
    const int pid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pid >= n_paths) return;
    const int off = path_off[pid], P = path_off[pid + 1] - off;
    const double d = offset[pid];
    const double2* g = pxy + off;
    for (int j = lane; j < P; j += 32) {
        double2 q = g[j];
        if (P >= 2) {
            const int k = (j == P - 1) ? P - 2 : j;
            const double2 n = dp_normal(g[k], g[k + 1]);
            q.x = fma(d, n.x, q.x); q.y = fma(d, n.y, q.y);
        }
        out_xy[off + j] = q;
    }
}

// CShare::BezierPlanning, one warp per pose pair
__global__ void __launch_bounds__(DP_WARPS_PER_BLOCK * 32)
op_bezier_kernel(int n, const double* __restrict__ poses, double* __restrict__ out_xy) {
    __shared__ WarpSmem smem[DP_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pid = blockIdx.x * DP_WARPS_PER_BLOCK + wib;
    if (pid >= n) return;
    const double* q = poses + (size_t)pid * 6;
    dp_bezier_to_plan(smem[wib], q[0], q[1], q[2], q[3], q[4], q[5], lane);
    for (int i = lane; i < DP_PATH_POINTS; i += 32) {
        out_xy[(size_t)pid * 400 + i] = smem[wib].plan[i].x;
        out_xy[(size_t)pid * 400 + DP_PATH_POINTS + i] = smem[wib].plan[i].y;
    }
}

// CShare::MeanPoints, one warp per path (n_in <= DP_SCR checked by the host)
__global__ void __launch_bounds__(DP_WARPS_PER_BLOCK * 32)
op_mean_kernel(int n_paths, const int32_t* __restrict__ path_off, const double2* __restrict__ pxy, double* __restrict__ out_xy) {
    __shared__ WarpSmem smem[DP_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pid = blockIdx.x * DP_WARPS_PER_BLOCK + wib;
    if (pid >= n_paths) return;
    const int off = path_off[pid], P = path_off[pid + 1] - off;
    dp_mean_points_to_plan(smem[wib], dp_src_run(pxy + off, 1, P), P, lane);
    for (int i = lane; i < DP_PATH_POINTS; i += 32) {
        out_xy[(size_t)pid * 400 + i] = smem[wib].plan[i].x;
        out_xy[(size_t)pid * 400 + DP_PATH_POINTS + i] = smem[wib].plan[i].y;
    }
}

// CShare::NearestId (call sites Decision.cpp:1889,2074,2383; loop shape Planning.cpp:640-648), one warp per (path, query)
__global__ void __launch_bounds__(DP_WARPS_PER_BLOCK * 32)
op_nearest_kernel(int n_paths, const int32_t* __restrict__ path_off, const double2* __restrict__ pxy, const double* __restrict__ qx,
                  const double* __restrict__ qy, int32_t* __restrict__ out_id) {
    const int lane = threadIdx.x & 31, pid = blockIdx.x * DP_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (pid >= n_paths) return;
    const int off = path_off[pid], P = path_off[pid + 1] - off;
    const int id = dp_nearest_plain(pxy + off, P, qx[pid], qy[pid], lane);
    if (lane == 0) out_id[pid] = (id == 0x7fffffff) ? 0 : id;   // the operator returns 0 when nothing is closer than 9999
}

// ------------------------------------------------------------------------------------------------
// Dense candidate sweep (BASELINE config 3).  A candidate = CreateNewPath(first P points of the base line, offset), scored by
// SearchObstacle against obstacle tracks o(j) = o0 + j dv.  Candidates that share an offset ("row") share their geometry:
// the candidate with P points consists of the row points q_j = b_j + off n_j, j < P-1, plus ONE own last point
// b_{P-1} + off n_{P-2} (CreateNewPath gives the last point the normal of the last segment, oracle/cshare_spec.h).  So the
// nearest-point search of every horizon of a row is ONE pass over the row per obstacle: the running argmin over q_0..q_{P-2}
// is the state, each horizon adds its own last point, runs the gates / lateral offset / corridor test with its own P, and
// min-reduces its packed (path index << 16 | obstacle) key; the arclength to the selected point is a look-up in the row's
// sequential prefix of segment lengths (every horizon starts at q_0, so the prefix IS the reference's sum).  Same bits as
// scoring each candidate alone (tests/test_gpu_parity.py::test_dense_sweep*), rows x obstacles x points evaluations instead
// of candidates x obstacles x points: 0.8 M instead of 250 M on the 64 x 32 x 32 grid.
//   sweep_rows_kernel   one thread-block CLUSTER per row, one CTA per block of 32 obstacles, a warp per part of the row: every
//                       lane of a warp walks the same row points and meets the same horizons at the same step (no divergence,
//                       row points broadcast); pass A: argmin of the part; pass B: the part again from the state the earlier
//                       parts left.  The CTAs of a row min-reduce their packed keys into the shared memory of the cluster's
//                       first CTA (distributed shared memory, remote atomicMin), which then owns the row's result.
//                       FUSED (latency session, dp_sweep_score): that CTA also selects among the row's horizon groups, the
//                       grid's last CTA writes the winner into page-locked host memory and re-arms the state -- one launch
//                       per call, obstacles in the kernel parameters, no copy / memset / graph nodes
//   sweep_prefix_kernel one CTA per row: sequential prefix of the row's segment lengths (obstacle-independent: once per candidate set)
//   sweep_select_kernel (dp_score_candidates only) one thread per candidate: cost of its (row, horizon) group, lowest feasible
//                       index by a packed (cost bits << 32 | index) 64-bit atomicMin: the reference's first-feasible `break`
//                       (Decision.cpp:944-953)
// ------------------------------------------------------------------------------------------------
#define SWEEP_MAX_BASE 256
#define SWEEP_MAX_OBS 192
#define SWEEP_MAXW 32                                       // warps (= parts of the row) per CTA at most

struct SweepObs { double v[4][SWEEP_MAX_OBS]; };            // x0, y0, dvx, dvy: travels in the kernel parameters (6 KB)
struct SweepSmem {
    double2 b[SWEEP_MAX_BASE];                              // base line
    double2 n[SWEEP_MAX_BASE];                              // unit right normal of segment j -> j+1
    double2 q[SWEEP_MAX_BASE];                              // row points b_j + off n_j
    unsigned key[SWEEP_MAX_BASE];                           // per horizon group of the row (first CTA of the cluster: the row's)
    int gP[SWEEP_MAX_BASE];                                 // the row's horizons (point counts), ascending
    double obs[4][32];                                      // this CTA's obstacle block
    unsigned long long red[SWEEP_MAXW];
    double redd[SWEEP_MAXW];
    int last;
    // followed (dynamic shared memory) by the running-argmin state at the end of each part of the row, per obstacle:
    // double pd[PT][32]; int pj[PT][32]
};
static size_t sweep_smem_bytes(int PT) { return sizeof(SweepSmem) + (size_t)PT * 32 * (sizeof(double) + sizeof(int)); }

// one row against this CTA's 32 obstacles (o0 ..): min-reduces the packed selection key of every horizon group of the row into
// rkey[0..ng) (the row owner's shared memory).  warp = part of the row; obs(k, o) = component k of obstacle track o
template <class ObsFn>
__device__ __forceinline__ void sweep_row_pass(SweepSmem& s, unsigned* rkey, double off, const int32_t* __restrict__ gP, int ng, int Pmax,
                                               int n_obs, int o0, int PT, double lat_min, double lat_max, ObsFn obs) {
    namespace cg = cooperative_groups;
    const int tid = threadIdx.x, nthr = blockDim.x, qt = tid >> 5, lane = tid & 31;
    double* s_pd = reinterpret_cast<double*>(&s + 1);
    int* s_pj = reinterpret_cast<int*>(s_pd + PT * 32);
    for (int g = tid; g < ng; g += nthr) { s.key[g] = 0xffffffffu; s.gP[g] = gP[g]; }
    for (int t = tid; t < 128; t += nthr) { const int k = t >> 5, o = o0 + (t & 31); s.obs[k][t & 31] = (o < n_obs) ? obs(k, o) : 0.0; }
    cg::this_cluster().sync();                              // (the owner's keys are armed before any CTA of the row reduces into them)
    for (int j = tid; j + 1 < Pmax; j += nthr) {
        const double2 n = dp_normal(s.b[j], s.b[j + 1]);
        s.n[j] = n;
        s.q[j] = make_double2(fma(off, n.x, s.b[j].x), fma(off, n.y, s.b[j].y));
    }
    __syncthreads();
    const int o = o0 + lane;
    const bool act = qt < PT && o < n_obs;
    const int nrow = Pmax > 0 ? Pmax - 1 : 0;               // row points q_0 .. q_{Pmax-2}
    const int qlen = (nrow + PT - 1) / PT, j0 = qt * qlen, j1 = min(nrow, j0 + qlen);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const double x0 = s.obs[0][lane], y0 = s.obs[1][lane], vx = s.obs[2][lane], vy = s.obs[3][lane];
    if (qt < PT && PT > 1) {   // pass A: argmin of the own part (one part: pass B alone is the whole scan)
        double bd = INF; int bj = 0;
        if (act) {
            double jd = (double)j0;
            for (int j = j0; j < j1; ++j, jd += 1.0) {
                const double2 q = s.q[j];
                const double dx = fma(jd, vx, x0) - q.x, dy = fma(jd, vy, y0) - q.y;
                const double d = fma(dx, dx, dy * dy);
                if (d < bd) { bd = d; bj = j; }
            }
        }
        s_pd[qt * 32 + lane] = bd; s_pj[qt * 32 + lane] = bj;
    }
    if (PT > 1) __syncthreads();
    if (act) {
        // pass B: start from the state at the end of the parts before mine (index order, strict '<': lowest index on ties),
        // walk my part again and serve the horizons whose last row point falls into it
        double bd = INF; int bj = 0;
        for (int q = 0; q < qt; ++q) {
            const double od = s_pd[q * 32 + lane];
            if (od < bd) { bd = od; bj = s_pj[q * 32 + lane]; }
        }
        // horizon P needs the state over q_0 .. q_{P-2}: it is served by the part that holds row point P-2 ... i.e. when the
        // walk is about to add row point j = P-1 (or, for the last horizon, at the end of the last part)
        int g = 0;
        for (int hi = ng; g < hi;) {                        // first horizon with P-1 >= j0 (earlier ones are served by earlier parts)
            const int mid = (g + hi) >> 1;
            if (s.gP[mid] - 1 < j0) g = mid + 1; else hi = mid;
        }
        int Pn = (g < ng) ? s.gP[g] : 0;
        const int jend = (qt == PT - 1 || j1 >= nrow) ? nrow + 1 : j1;   // the last part also serves P = Pmax (after the last row point)
        double jd = (double)j0;
        for (int j = j0; j < jend; ++j, jd += 1.0) {
            const double mx = fma(jd, vx, x0), my = fma(jd, vy, y0);   // obstacle when the ego reaches path point j
            while (g < ng && Pn == j + 1) {
                // horizon P = j + 1: its points are q_0 .. q_{P-2} (state bd, bj) and its own last point b_{P-1} + off n_{P-2}
                const int P = Pn;
                const double2 nl = s.n[P - 2], bl = s.b[P - 1];
                const double2 ql = make_double2(fma(off, nl.x, bl.x), fma(off, nl.y, bl.y));
                const double dx = mx - ql.x, dy = my - ql.y;
                const double dl = fma(dx, dx, dy * dy);
                const int cj = (dl < bd) ? P - 1 : bj;      // strict '<': the last index only wins when strictly closer
                const double hx = fma((double)cj, vx, x0), hy = fma((double)cj, vy, y0);   // the obstacle at step cj
                const int k = (cj == P - 1) ? P - 2 : cj;
                const double2 pk = s.q[k], pk1 = (k + 1 == P - 1) ? ql : s.q[k + 1];
                double dd;
                unsigned key = dp_owner_key(pk, pk1, cj, P, o, hx, hy, lat_min, lat_max, &dd);
                const unsigned am = __activemask();         // (the lanes that arrive together, normally all active ones: one atomic for them)
                key = __reduce_min_sync(am, key);
                if (key != 0xffffffffu && lane == __ffs(am) - 1) atomicMin(&rkey[g], key);
                ++g;
                Pn = (g < ng) ? s.gP[g] : 0;
            }
            if (j < j1) {                                   // row point j enters the running argmin (it is point j of every longer horizon)
                const double2 q = s.q[j];
                const double dx = mx - q.x, dy = my - q.y;
                const double d = fma(dx, dx, dy * dy);
                if (d < bd) { bd = d; bj = j; }
            }
        }
    }
    cg::this_cluster().sync();                              // every CTA's keys have landed in the owner's shared memory
}

// arclength of horizon group g of the row to its selected point = the row prefix (+ one term when the selected point is the
// horizon's own last point b_{P-1} + off n_{P-2}: q_{P-2} takes the normal of segment P-2 too)
__device__ __forceinline__ double sweep_group_dis(const SweepSmem& s, int g, double off, const double* __restrict__ cum) {
    const unsigned k = s.key[g];
    if (k == 0xffffffffu) return DP_NOT_FOUND;              // nothing in the corridor
    const int P = s.gP[g], jstar = (int)(k >> 16);
    if (jstar <= P - 2) return cum[jstar];
    const double2 nl = s.n[P - 2], b1 = s.b[P - 1], b2 = s.b[P - 2];
    const double2 ql = make_double2(fma(off, nl.x, b1.x), fma(off, nl.y, b1.y));
    const double2 qp = make_double2(fma(off, nl.x, b2.x), fma(off, nl.y, b2.y));
    return cum[P - 2] + sqrt(dp_sq2(ql.x - qp.x, ql.y - qp.y));
}

struct SweepOut {                                           // FUSED: where the winner goes and the state the last CTA re-arms
    unsigned long long* row_res;                            // device: {packed key, dis_lng bits} of every row's best group
    unsigned* done;                                         // device: rows finished (0 between calls)
    const int32_t* group_first;                             // lowest candidate index of each group
    int first_nogroup;                                      // lowest candidate with fewer than 2 points (-1: none): dis_lng = DP_NOT_FOUND
    volatile unsigned long long* host;                      // page-locked: {packed key, dis_lng bits, sequence number}
    unsigned long long seq;
    long long* dbg;                                         // diagnostic (DP_SWEEP_DBG=1): globaltimer stamps [row][8] of the row owners
};
#define SWEEP_STAMP(k) if (FUSED && out.dbg && threadIdx.x == 0 && ob == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); out.dbg[(size_t)(blockIdx.x / NB) * 8 + (k)] = (long long)t_; }

// grid = max(n_rows, 1) clusters of NB CTAs (cluster dimension set by the launcher): CTA = (row blockIdx.x / NB, obstacle block blockIdx.x % NB)
template <bool FUSED>
__global__ void __launch_bounds__(1024)
sweep_rows_kernel(const double* __restrict__ lines, int n_base, int n_lines, int n_rows, const double* __restrict__ row_off,
                  const int4* __restrict__ row_info, const int32_t* __restrict__ group_P, const double* __restrict__ ox,
                  const double* __restrict__ oy, const double* __restrict__ dvx, const double* __restrict__ dvy,
                  const __grid_constant__ SweepObs po, int n_obs, int NB, int PT, double lat_min, double lat_max, double clear_dis,
                  unsigned* __restrict__ group_key, const double* __restrict__ row_cum, SweepOut out) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char sweep_raw[];
    SweepSmem& s = *reinterpret_cast<SweepSmem*>(sweep_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int ob = (int)cluster.block_rank(), row = blockIdx.x / NB, tid = threadIdx.x, nthr = blockDim.x;
    unsigned long long mykey = ~0ull;
    double mydis = DP_NOT_FOUND;
    SWEEP_STAMP(0);
    if (row < n_rows) {                                     // (uniform over the cluster)
        // row header {first group, groups, longest horizon, base line}.  With ONE base line its loads do not depend on the header
        // and fly together with it
        const int4 ri = row_info[row];
        const double off = row_off[row];
        const double* bx = lines + (n_lines > 1 ? (size_t)ri.w * 2 * n_base : 0);
        for (int j = tid; j < n_base; j += nthr) s.b[j] = make_double2(bx[j], bx[n_base + j]);
        const int g0 = ri.x, ng = ri.y, Pmax = ri.z;
        unsigned* rkey = cluster.map_shared_rank(s.key, 0);
        if (FUSED)
            sweep_row_pass(s, rkey, off, group_P + g0, ng, Pmax, n_obs, ob * 32, PT, lat_min, lat_max,
                           [&](int k, int o) { return po.v[k][o]; });
        else
            sweep_row_pass(s, rkey, off, group_P + g0, ng, Pmax, n_obs, ob * 32, PT, lat_min, lat_max,
                           [&](int k, int o) { return k == 0 ? ox[o] : k == 1 ? oy[o] : k == 2 ? (dvx ? dvx[o] : 0.0) : (dvy ? dvy[o] : 0.0); });
        SWEEP_STAMP(1);
        if (ob != 0) return;                                // the row's first CTA carries on with the merged keys
        if (!FUSED) {
            for (int g = tid; g < ng; g += nthr) group_key[g0 + g] = s.key[g];
            return;
        }
        const double* cum = row_cum + (size_t)row * SWEEP_MAX_BASE;
        for (int g = tid; g < ng; g += nthr) {
            const double dis = sweep_group_dis(s, g, off, cum);
            const float cost = (dis > clear_dis) ? 0.0f : __int_as_float(0x7f800000);
            const unsigned long long k = ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)out.group_first[g0 + g];
            if (k < mykey) { mykey = k; mydis = dis; }
        }
    }
    if (!FUSED || ob != 0) return;
    SWEEP_STAMP(2);
    if (blockIdx.x == 0 && tid == 0 && out.first_nogroup >= 0) {
        const float cost = (DP_NOT_FOUND > clear_dis) ? 0.0f : __int_as_float(0x7f800000);
        const unsigned long long k = ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)out.first_nogroup;
        if (k < mykey) { mykey = k; mydis = DP_NOT_FOUND; }
    }
    // keys are distinct (they end in a candidate index) unless both are ~0: the dis_lng travels with the smaller key
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(DP_FULL, mykey, o);
        const double od = __shfl_xor_sync(DP_FULL, mydis, o);
        if (ok < mykey) { mykey = ok; mydis = od; }
    }
    if ((tid & 31) == 0) { s.red[tid >> 5] = mykey; s.redd[tid >> 5] = mydis; }
    if (tid == 0) s.last = 0;
    __syncthreads();
    if (tid >= 32) return;
    mykey = (tid < (nthr >> 5)) ? s.red[tid] : ~0ull;
    mydis = (tid < (nthr >> 5)) ? s.redd[tid] : DP_NOT_FOUND;
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(DP_FULL, mykey, o);
        const double od = __shfl_xor_sync(DP_FULL, mydis, o);
        if (ok < mykey) { mykey = ok; mydis = od; }
    }
    SWEEP_STAMP(3);
    const int rows = n_rows > 0 ? n_rows : 1;
    int last = 0;
    if (tid == 0) {
        const int r = row < rows ? row : 0;
        out.row_res[2 * r] = mykey; out.row_res[2 * r + 1] = (unsigned long long)__double_as_longlong(mydis);
        __threadfence();
        last = atomicAdd(out.done, 1u) == (unsigned)rows - 1u;
        if (last) { *out.done = 0; __threadfence(); }       // re-armed for the next call (stream order)
    }
    last = __shfl_sync(DP_FULL, last, 0);
    SWEEP_STAMP(4);
    if (!last) return;
    // the last row: first warp reduces the rows' results and publishes the winner
    mykey = ~0ull; mydis = DP_NOT_FOUND;
    for (int r = tid; r < rows; r += 32) {
        const unsigned long long k = __ldcg(out.row_res + 2 * r);
        const unsigned long long d = __ldcg(out.row_res + 2 * r + 1);
        if (k < mykey) { mykey = k; mydis = __longlong_as_double((long long)d); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(DP_FULL, mykey, o);
        const double od = __shfl_xor_sync(DP_FULL, mydis, o);
        if (ok < mykey) { mykey = ok; mydis = od; }
    }
    if (tid == 0) {
        SWEEP_STAMP(5);
        out.host[0] = mykey; out.host[1] = (unsigned long long)__double_as_longlong(mydis);
        __threadfence_system();
        SWEEP_STAMP(6);
        out.host[2] = out.seq;
    }
}

// the row's sequential prefix of segment lengths (all horizons start at q_0, so the prefix IS the reference's arclength sum):
// row_cum[row][j] = |q_1 - q_0| + ... + |q_j - q_{j-1}|, j <= Pmax - 2.  Independent of the obstacles: once per candidate set.
__global__ void __launch_bounds__(256)
sweep_prefix_kernel(const double* __restrict__ lines, int n_base, const double* __restrict__ row_off, const int4* __restrict__ row_info,
                    double* __restrict__ row_cum) {
    __shared__ double2 s_b[SWEEP_MAX_BASE];
    __shared__ double2 s_q[SWEEP_MAX_BASE];
    __shared__ double s_cum[SWEEP_MAX_BASE];
    const int row = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int4 ri = row_info[row];
    const double off = row_off[row];
    const int Pmax = ri.z;
    const double* bx = lines + (size_t)ri.w * 2 * n_base;
    for (int j = tid; j < Pmax; j += nthr) s_b[j] = make_double2(bx[j], bx[n_base + j]);
    __syncthreads();
    for (int j = tid; j + 1 < Pmax; j += nthr) {
        const double2 n = dp_normal(s_b[j], s_b[j + 1]);
        s_q[j] = make_double2(fma(off, n.x, s_b[j].x), fma(off, n.y, s_b[j].y));
    }
    __syncthreads();
    for (int j = tid; j + 2 < Pmax; j += nthr) s_cum[j + 1] = sqrt(dp_sq2(s_q[j + 1].x - s_q[j].x, s_q[j + 1].y - s_q[j].y));
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        s_cum[0] = 0.0;
        for (int j = 1; j + 1 < Pmax; ++j) { acc += s_cum[j]; s_cum[j] = acc; }
    }
    __syncthreads();
    for (int j = tid; j + 1 < Pmax; j += nthr) row_cum[(size_t)row * SWEEP_MAX_BASE + j] = s_cum[j];
}

// one thread per candidate: arclength of its (row, horizon) group to the selected point = the row prefix (+ one term when the
// selected point is the horizon's own last point), cost, lowest feasible index by a packed (cost bits << 32 | index) atomicMin
__global__ void sweep_select_kernel(const int32_t* __restrict__ cand_group, int n_cand, const int32_t* __restrict__ group_row,
                                    const int32_t* __restrict__ group_P, const unsigned* __restrict__ group_key, const double* __restrict__ row_cum,
                                    const double* __restrict__ row_off, const double* __restrict__ lines, int n_base,
                                    const int4* __restrict__ row_info, double clear_dis, double* __restrict__ cand_dis_lng,
                                    unsigned long long* __restrict__ best_key) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = ~0ull;
    if (c < n_cand) {
        const int g = cand_group[c];
        double dis = DP_NOT_FOUND;                          // fewer than 2 points, or nothing in the corridor
        const unsigned k = g >= 0 ? group_key[g] : 0xffffffffu;
        if (k != 0xffffffffu) {
            const int row = group_row[g], P = group_P[g], jstar = (int)(k >> 16);
            const double* cum = row_cum + (size_t)row * SWEEP_MAX_BASE;
            if (jstar <= P - 2) dis = cum[jstar];
            else {                                          // the horizon's own last point b_{P-1} + off n_{P-2}: one more term after q_{P-2}
                const double off = row_off[row];
                const double* bx = lines + (size_t)row_info[row].w * 2 * n_base;
                const double2 b2 = make_double2(bx[P - 2], bx[n_base + P - 2]), b1 = make_double2(bx[P - 1], bx[n_base + P - 1]);
                const double2 nl = dp_normal(b2, b1);
                const double2 ql = make_double2(fma(off, nl.x, b1.x), fma(off, nl.y, b1.y));
                const double2 qp = make_double2(fma(off, nl.x, b2.x), fma(off, nl.y, b2.y));   // q_{P-2} uses the normal of segment P-2 too
                dis = cum[P - 2] + sqrt(dp_sq2(ql.x - qp.x, ql.y - qp.y));
            }
        }
        if (cand_dis_lng) cand_dis_lng[c] = dis;
        const float cost = (dis > clear_dis) ? 0.0f : __int_as_float(0x7f800000);
        key = ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)c;
    }
    for (int o = 16; o > 0; o >>= 1) {                      // lowest key of the warp, one atomic per warp
        const unsigned long long other = __shfl_xor_sync(DP_FULL, key, o);
        key = other < key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key != ~0ull) atomicMin(best_key, key);
}

// FMA micro-benchmark: 8 independent accumulator chains per thread
template <typename T>
__global__ void fma_peak_kernel(T* sink, int iters) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3, a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6,
      a7 = a0 + (T)7;
    const T m = (T)0.999999, b = (T)1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)123456789) sink[0] = s;                    // never true: keeps the chains alive
}

}  // namespace

cudaError_t dp_launch_search(int n_paths, const int32_t* path_off, const double2* pxy, const double* ox,
                             const double* oy, int n_obs, const double* lat_min, const double* lat_max, dp_search_slot* out,
                             cudaStream_t st) {
    if (n_paths <= 0) return cudaSuccess;
    op_search_kernel<<<(n_paths + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK, DP_WARPS_PER_BLOCK * 32, 0, st>>>(
        n_paths, path_off, pxy, ox, oy, n_obs, lat_min, lat_max, out);
    return cudaGetLastError();
}
cudaError_t dp_launch_create(int n_paths, const int32_t* path_off, const double2* pxy, const double* offset,
                             double2* out_xy, cudaStream_t st) {
    if (n_paths <= 0) return cudaSuccess;
    op_create_kernel<<<(n_paths * 32 + 127) / 128, 128, 0, st>>>(n_paths, path_off, pxy, offset, out_xy);
    return cudaGetLastError();
}
cudaError_t dp_launch_bezier(int n, const double* poses, double* out_xy, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    op_bezier_kernel<<<(n + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK, DP_WARPS_PER_BLOCK * 32, 0, st>>>(n, poses, out_xy);
    return cudaGetLastError();
}
cudaError_t dp_launch_mean(int n_paths, const int32_t* path_off, const double2* pxy, double* out_xy, cudaStream_t st) {
    if (n_paths <= 0) return cudaSuccess;
    op_mean_kernel<<<(n_paths + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK, DP_WARPS_PER_BLOCK * 32, 0, st>>>(n_paths, path_off, pxy, out_xy);
    return cudaGetLastError();
}
cudaError_t dp_launch_nearest(int n_paths, const int32_t* path_off, const double2* pxy, const double* qx, const double* qy, int32_t* out_id,
                              cudaStream_t st) {
    if (n_paths <= 0) return cudaSuccess;
    op_nearest_kernel<<<(n_paths + DP_WARPS_PER_BLOCK - 1) / DP_WARPS_PER_BLOCK, DP_WARPS_PER_BLOCK * 32, 0, st>>>(n_paths, path_off, pxy, qx, qy, out_id);
    return cudaGetLastError();
}
// the obstacle-independent part of a sweep: once per candidate set
cudaError_t dp_launch_sweep_prefix(const double* lines, int n_base, int n_rows, const double* row_off, const int4* row_info, double* row_cum,
                                   cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    sweep_prefix_kernel<<<n_rows, 256, 0, st>>>(lines, n_base, row_off, row_info, row_cum);
    return cudaGetLastError();
}
// shape of the row clusters: NB CTAs (obstacle blocks of 32) x PT warps (parts of the row).  Few rows (latency mode, the machine
// is mostly idle): 16 parts shorten the dependent chain of a row; many rows (more than two per SM): 4 parts -- measured on the
// 2048-row Bezier grid of bench.py: 197 / 144 / 103 / 117 / 158 us for 1 / 2 / 4 / 8 / 16 parts (fewer parts leave the SM short of
// warps because every CTA carries the row's 14 KB of points, more parts repeat the per-part overhead)
static void sweep_shape(int n_rows, int n_obs, int* NB, int* PT) {
    *NB = n_obs > 0 ? (n_obs + 31) / 32 : 1;
    int pt = n_rows > 296 ? 4 : 16;
    if (const char* e = getenv("DP_SWEEP_PARTS")) pt = atoi(e);
    *PT = pt < 1 ? 1 : (pt > SWEEP_MAXW ? SWEEP_MAXW : pt);
}
template <bool FUSED, class... Args>
static cudaError_t sweep_launch(int n_rows, int NB, int PT, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n_rows > 0 ? n_rows : 1) * NB)); cfg.blockDim = dim3((unsigned)(PT * 32)); cfg.stream = st;
    cfg.dynamicSmemBytes = sweep_smem_bytes(PT);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)NB; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, sweep_rows_kernel<FUSED>, args...);
}
cudaError_t dp_launch_sweep(const double* lines, int n_base, int n_lines, int n_rows, const double* row_off, const int4* row_info,
                            const int32_t* group_P, const int32_t* group_row, const int32_t* cand_group, int n_cand, const double* ox,
                            const double* oy, const double* dvx, const double* dvy, int n_obs, double lat_min, double lat_max, double clear_dis,
                            unsigned* group_key, const double* row_cum, double* cand_dis_lng, unsigned long long* best_key, cudaStream_t st) {
    if (n_cand <= 0) return cudaSuccess;
    if (n_rows > 0) {
        int NB, PT;
        sweep_shape(n_rows, n_obs, &NB, &PT);
        static const SweepObs none = {};
        const cudaError_t e = sweep_launch<false>(n_rows, NB, PT, st, lines, n_base, n_lines, n_rows, row_off, row_info, group_P, ox, oy, dvx, dvy, none,
                                                  n_obs, NB, PT, lat_min, lat_max, clear_dis, group_key, row_cum, SweepOut{});
        if (e != cudaSuccess) return e;
    }
    sweep_select_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>(cand_group, n_cand, group_row, group_P, group_key, row_cum, row_off, lines, n_base,
                                                              row_info, clear_dis, cand_dis_lng, best_key);
    return cudaGetLastError();
}
// latency session: ONE launch; the winner lands in host[0..2] (page-locked), host[2] = seq last
cudaError_t dp_launch_sweep_fused(const double* lines, int n_base, int n_lines, int n_rows, const double* row_off, const int4* row_info,
                                  const int32_t* group_P, const int32_t* group_first, int first_nogroup, const double* obs4, int obs_stride,
                                  int n_obs, double lat_min, double lat_max, double clear_dis, const double* row_cum,
                                  unsigned long long* row_res, unsigned* done, unsigned long long* host, unsigned long long seq, long long* dbg,
                                  cudaStream_t st) {
    SweepObs po;                                            // (copied into the launch: no lifetime beyond the call)
    for (int k = 0; k < 4; ++k) memcpy(po.v[k], obs4 + (size_t)k * obs_stride, (size_t)n_obs * sizeof(double));
    int NB, PT;
    sweep_shape(n_rows, n_obs, &NB, &PT);
    SweepOut out;
    out.row_res = row_res; out.done = done; out.group_first = group_first; out.first_nogroup = first_nogroup; out.host = host; out.seq = seq;
    out.dbg = dbg;
    const double* nul = nullptr; unsigned* nkey = nullptr;
    return sweep_launch<true>(n_rows, NB, PT, st, lines, n_base, n_lines, n_rows, row_off, row_info, group_P, nul, nul, nul, nul, po, n_obs, NB, PT,
                              lat_min, lat_max, clear_dis, nkey, row_cum, out);
}
cudaError_t dp_launch_fma_peak(int which, float* sink, int iters, int blocks, cudaStream_t st) {
    if (which == 0) fma_peak_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<double*>(sink), iters);
    else fma_peak_kernel<float><<<blocks, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}
