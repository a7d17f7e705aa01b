// csrc/dp_ops.cu -- operator-level kernels behind the CShare seam, the
// dense candidate sweep of BASELINE config 3, and the FMA micro-benchmark that provides the
// roofline denominator (MEASURED_PEAKS.json has no FP64/FP32 CUDA-core figure).
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
//   sweep_rows_kernel   one CTA per (row, block of 32 obstacles): 8 threads per obstacle, each owns an eighth of the row
//                       (pass A: argmin of the part; pass B: the part again from the state the earlier parts left)
//   sweep_prefix_kernel one CTA per row: sequential prefix of the row's segment lengths (obstacle-independent: once per candidate set)
//   sweep_select_kernel one thread per candidate: cost of its (row, horizon) group, lowest feasible index by a packed
//                       (cost bits << 32 | index) 64-bit atomicMin: the reference's first-feasible `break` (Decision.cpp:944-953)
// ------------------------------------------------------------------------------------------------
#define SWEEP_MAX_BASE 256
#define SWEEP_MAX_OBS 192

#define SWEEP_Q 8                                           // threads per (row, obstacle): eighths of the row
#define SWEEP_OB 32                                         // obstacles per CTA
__global__ void __launch_bounds__(SWEEP_Q * SWEEP_OB)
sweep_rows_kernel(const double* __restrict__ base_x, const double* __restrict__ base_y, int n_base, const double* __restrict__ row_off,
                  const int32_t* __restrict__ row_gbeg, const int32_t* __restrict__ group_P, const double* __restrict__ ox,
                  const double* __restrict__ oy, const double* __restrict__ dvx, const double* __restrict__ dvy, int n_obs, int ob0, double lat_min,
                  double lat_max, unsigned* __restrict__ group_key) {
    __shared__ double2 s_b[SWEEP_MAX_BASE];                 // base line
    __shared__ double2 s_n[SWEEP_MAX_BASE];                 // unit right normal of segment j -> j+1
    __shared__ double2 s_q[SWEEP_MAX_BASE];                 // row points b_j + off n_j
    __shared__ unsigned s_key[SWEEP_MAX_BASE];              // per horizon group of the row
    __shared__ double s_pd[SWEEP_OB][SWEEP_Q];              // running-argmin state at the end of each part of the row, per obstacle
    __shared__ int s_pj[SWEEP_OB][SWEEP_Q];
    __shared__ int s_gP[SWEEP_MAX_BASE];                    // the row's horizons (point counts), ascending
    const int row = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int g0 = row_gbeg[row], ng = row_gbeg[row + 1] - g0;
    const double off = row_off[row];
    const int Pmax = ng > 0 ? group_P[g0 + ng - 1] : 0;     // horizons are sorted ascending
    for (int j = tid; j < Pmax; j += nthr) s_b[j] = make_double2(base_x[j], base_y[j]);
    for (int g = tid; g < ng; g += nthr) { s_key[g] = 0xffffffffu; s_gP[g] = group_P[g0 + g]; }
    __syncthreads();
    for (int j = tid; j + 1 < Pmax; j += nthr) {
        const double2 n = dp_normal(s_b[j], s_b[j + 1]);
        s_n[j] = n;
        s_q[j] = make_double2(fma(off, n.x, s_b[j].x), fma(off, n.y, s_b[j].y));
    }
    __syncthreads();
    // thread (obstacle ol, part qt): row points [j0, j1) of the running argmin; the obstacle block of this CTA
    const int ol = tid / SWEEP_Q, qt = tid - ol * SWEEP_Q, o = ob0 + blockIdx.y * SWEEP_OB + ol;
    const bool act = tid < SWEEP_OB * SWEEP_Q && o < n_obs;
    const int nrow = Pmax > 0 ? Pmax - 1 : 0;               // row points q_0 .. q_{Pmax-2}
    const int qlen = (nrow + SWEEP_Q - 1) / SWEEP_Q, j0 = qt * qlen, j1 = min(nrow, j0 + qlen);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double x0 = 0, y0 = 0, vx = 0, vy = 0;
    if (act) { x0 = ox[o]; y0 = oy[o]; vx = dvx ? dvx[o] : 0.0; vy = dvy ? dvy[o] : 0.0; }
    {   // pass A: argmin of the own quarter
        double bd = INF; int bj = 0;
        if (act) {
            double jd = (double)j0;
            for (int j = j0; j < j1; ++j, jd += 1.0) {
                const double2 q = s_q[j];
                const double dx = fma(jd, vx, x0) - q.x, dy = fma(jd, vy, y0) - q.y;
                const double d = fma(dx, dx, dy * dy);
                if (d < bd) { bd = d; bj = j; }
            }
            s_pd[ol][qt] = bd; s_pj[ol][qt] = bj;
        }
    }
    __syncthreads();
    if (act) {
        // pass B: start from the state at the end of the quarters before mine (index order, strict '<': lowest index on ties),
        // walk my quarter again and serve the horizons whose last row point falls into it
        double bd = INF; int bj = 0;
        for (int q = 0; q < qt; ++q) if (s_pd[ol][q] < bd) { bd = s_pd[ol][q]; bj = s_pj[ol][q]; }
        // horizon P needs the state over q_0 .. q_{P-2}: it is served by the quarter that holds row point P-2 ... i.e. when the
        // walk is about to add row point j = P-1 (or, for the last horizon, at the end of the last quarter)
        int g = 0;
        while (g < ng && s_gP[g] - 1 < j0) ++g;             // horizons served by earlier parts (P-1 < j0); P-1 == j0 is mine
        int Pn = (g < ng) ? s_gP[g] : 0;
        const int jend = (qt == SWEEP_Q - 1 || j1 >= nrow) ? nrow + 1 : j1;   // the last quarter also serves P = Pmax (after the last row point)
        double jd = (double)j0;
        for (int j = j0; j < jend; ++j, jd += 1.0) {
            const double mx = fma(jd, vx, x0), my = fma(jd, vy, y0);   // obstacle when the ego reaches path point j
            while (g < ng && Pn == j + 1) {
                // horizon P = j + 1: its points are q_0 .. q_{P-2} (state bd, bj) and its own last point b_{P-1} + off n_{P-2}
                const int P = Pn;
                const double2 nl = s_n[P - 2], bl = s_b[P - 1];
                const double2 ql = make_double2(fma(off, nl.x, bl.x), fma(off, nl.y, bl.y));
                const double dx = mx - ql.x, dy = my - ql.y;
                const double dl = fma(dx, dx, dy * dy);
                const int cj = (dl < bd) ? P - 1 : bj;      // strict '<': the last index only wins when strictly closer
                const double hx = fma((double)cj, vx, x0), hy = fma((double)cj, vy, y0);   // the obstacle at step cj
                const int k = (cj == P - 1) ? P - 2 : cj;
                const double2 pk = s_q[k], pk1 = (k + 1 == P - 1) ? ql : s_q[k + 1];
                double dd;
                const unsigned key = dp_owner_key(pk, pk1, cj, P, o, hx, hy, lat_min, lat_max, &dd);
                if (key != 0xffffffffu) atomicMin(&s_key[g], key);
                ++g;
                Pn = (g < ng) ? s_gP[g] : 0;
            }
            if (j < j1) {                                   // row point j enters the running argmin (it is point j of every longer horizon)
                const double2 q = s_q[j];
                const double dx = mx - q.x, dy = my - q.y;
                const double d = fma(dx, dx, dy * dy);
                if (d < bd) { bd = d; bj = j; }
            }
        }
    }
    __syncthreads();
    for (int g = tid; g < ng; g += nthr) if (s_key[g] != 0xffffffffu) atomicMin(&group_key[g0 + g], s_key[g]);
}

// the row's sequential prefix of segment lengths (all horizons start at q_0, so the prefix IS the reference's arclength sum):
// row_cum[row][j] = |q_1 - q_0| + ... + |q_j - q_{j-1}|, j <= Pmax - 2.  Independent of the obstacles: once per candidate set.
__global__ void __launch_bounds__(256)
sweep_prefix_kernel(const double* __restrict__ base_x, const double* __restrict__ base_y, const double* __restrict__ row_off,
                    const int32_t* __restrict__ row_gbeg, const int32_t* __restrict__ group_P, double* __restrict__ row_cum) {
    __shared__ double2 s_b[SWEEP_MAX_BASE];
    __shared__ double2 s_q[SWEEP_MAX_BASE];
    __shared__ double s_cum[SWEEP_MAX_BASE];
    const int row = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int g0 = row_gbeg[row], ng = row_gbeg[row + 1] - g0;
    const double off = row_off[row];
    const int Pmax = ng > 0 ? group_P[g0 + ng - 1] : 0;
    for (int j = tid; j < Pmax; j += nthr) s_b[j] = make_double2(base_x[j], base_y[j]);
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
                                    const double* __restrict__ row_off, const double* __restrict__ base_x, const double* __restrict__ base_y,
                                    double clear_dis, double* __restrict__ cand_dis_lng, unsigned long long* __restrict__ best_key) {
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
                const double2 b2 = make_double2(base_x[P - 2], base_y[P - 2]), b1 = make_double2(base_x[P - 1], base_y[P - 1]);
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
cudaError_t dp_launch_sweep_prefix(const double* base_x, const double* base_y, int n_rows, const double* row_off, const int32_t* row_gbeg,
                                   const int32_t* group_P, double* row_cum, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    sweep_prefix_kernel<<<n_rows, 256, 0, st>>>(base_x, base_y, row_off, row_gbeg, group_P, row_cum);
    return cudaGetLastError();
}
cudaError_t dp_launch_sweep(const double* base_x, const double* base_y, int n_base, int n_rows, int n_groups, const double* row_off,
                            const int32_t* row_gbeg, const int32_t* group_P, const int32_t* group_row, const int32_t* cand_group, int n_cand,
                            const double* ox, const double* oy, const double* dvx, const double* dvy, int n_obs, double lat_min, double lat_max,
                            double clear_dis, unsigned* group_key, const double* row_cum, double* cand_dis_lng, unsigned long long* best_key,
                            cudaStream_t st) {
    if (n_cand <= 0) return cudaSuccess;
    if (n_rows > 0) {
        cudaError_t e = cudaMemsetAsync(group_key, 0xff, (size_t)n_groups * sizeof(unsigned), st);
        if (e != cudaSuccess) return e;
        if (n_obs > 0) {
            const dim3 grid(n_rows, (n_obs + SWEEP_OB - 1) / SWEEP_OB);   // a CTA = one row x 32 obstacles x 8 parts of the row
            sweep_rows_kernel<<<grid, SWEEP_Q * SWEEP_OB, 0, st>>>(base_x, base_y, n_base, row_off, row_gbeg, group_P, ox, oy, dvx, dvy, n_obs, 0,
                                                                  lat_min, lat_max, group_key);
        }
    }
    sweep_select_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>(cand_group, n_cand, group_row, group_P, group_key, row_cum, row_off, base_x, base_y,
                                                              clear_dis, cand_dis_lng, best_key);
    return cudaGetLastError();
}
cudaError_t dp_launch_fma_peak(int which, float* sink, int iters, int blocks, cudaStream_t st) {
    if (which == 0) fma_peak_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<double*>(sink), iters);
    else fma_peak_kernel<float><<<blocks, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}
