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
// Dense candidate sweep (BASELINE config 3): one warp per candidate, SWEEP_WARPS candidates per
// block.  The base polyline and its segment normals are staged once per block in shared memory,
// obstacle tracks (x, y, dvx, dvy) too; each warp materialises ITS candidate (offset copy of a
// prefix of the base line) in its own shared-memory slab, then lanes = obstacles run the fused
// rollout -> nearest-point -> corridor check; arclength is summed in index order.  Selection is
// a packed (cost bits << 32 | candidate index) 64-bit atomicMin: feasible candidates have cost 0,
// so the minimum is the LOWEST feasible index (the reference's first-feasible `break`,
// Decision.cpp:944-953), deterministic whatever the launch geometry.
// ------------------------------------------------------------------------------------------------
#define SWEEP_WARPS 8
#ifndef SWEEP_CTAS_PER_SM
#define SWEEP_CTAS_PER_SM 3
#endif
#define SWEEP_MAX_BASE 256
#define SWEEP_MAX_OBS 192

// argmin update (a 64-bit integer compare on the bit patterns was tried to unload the FP64 pipe: 2 ISETP + 3 SEL cost more
// issue slots than DSETP + 3 SEL, and this kernel is issue-bound -- profiles/README.md)
__device__ __forceinline__ void sweep_upd(double d2, int j, double& bb, int& bj) {
    if (d2 < bb) { bb = d2; bj = j; }
}

__global__ void __launch_bounds__(SWEEP_WARPS * 32, SWEEP_CTAS_PER_SM)
sweep_kernel(const double* __restrict__ base_x, const double* __restrict__ base_y, int n_base, const double* __restrict__ offset,
             const int32_t* __restrict__ n_pts, int n_cand, const double* __restrict__ ox, const double* __restrict__ oy,
             const double* __restrict__ dvx, const double* __restrict__ dvy, int n_obs, double lat_min, double lat_max,
             double clear_dis, double* __restrict__ cand_dis_lng, unsigned long long* __restrict__ best_key,
             const int32_t* __restrict__ order, unsigned* __restrict__ next) {
    __shared__ double2 s_base[SWEEP_MAX_BASE];
    __shared__ double2 s_nrm[SWEEP_MAX_BASE];              // normal of segment j -> j+1
    __shared__ double4 s_obs[SWEEP_MAX_OBS];
    __shared__ double2 s_cand[SWEEP_WARPS][SWEEP_MAX_BASE]; // the warp's candidate; reused for the arclength terms afterwards
    __shared__ unsigned long long s_key[SWEEP_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < n_base; j += blockDim.x) s_base[j] = make_double2(base_x[j], base_y[j]);
    for (int o = threadIdx.x; o < n_obs; o += blockDim.x) s_obs[o] = make_double4(ox[o], oy[o], dvx ? dvx[o] : 0.0, dvy ? dvy[o] : 0.0);
    __syncthreads();
    for (int j = threadIdx.x; j + 1 < n_base; j += blockDim.x) s_nrm[j] = dp_normal(s_base[j], s_base[j + 1]);
    __syncthreads();

    unsigned long long mykey = ~0ull;
    // Candidates differ 32x in length (8..256 points): a static round-robin leaves the slowest warp ~30 % behind the mean.
    // Warps fetch the next candidate from a counter instead, in the order the session prepared (longest first).
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(next, 1u);
        i = __shfl_sync(DP_FULL, i, 0);
        if (i >= n_cand) break;
        const int c = order ? order[i] : i;
        const int P = min(n_pts[c], n_base);
        const double off = offset[c];
        double dis_lng = DP_NOT_FOUND;
        if (P >= 2) {
            double2* q = s_cand[wib];
            __syncwarp();
            for (int j = lane; j < P; j += 32) {           // rollout: offset copy, never leaves the SM
                const double2 b = s_base[j], n = s_nrm[min(j, P - 2)];
                q[j] = make_double2(fma(off, n.x, b.x), fma(off, n.y, b.y));
            }
            __syncwarp();
            unsigned bestkey = 0xffffffffu;
            for (int g = 0; g * 32 < n_obs; ++g) {
                const int o = g * 32 + lane;
                if (o < n_obs) {
                    const double4 ob = s_obs[o];
                    // 4 independent running minima (j mod 4): lexicographic (d2, j) min == sequential strict-'<' argmin
                    const double INF = __longlong_as_double(0x7ff0000000000000LL);
                    double b0 = INF, b1 = INF, b2 = INF, b3 = INF;
                    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
                    int j = 0;
                    double jd = 0.0;                        // (double)j, kept as a running exact integer
                    for (; j + 4 <= P; j += 4, jd += 4.0) {
                        const double2 p0 = q[j], p1 = q[j + 1], p2 = q[j + 2], p3 = q[j + 3];
                        const double t0 = jd, t1 = jd + 1.0, t2 = jd + 2.0, t3 = jd + 3.0;
                        const double x0 = fma(t0, ob.z, ob.x) - p0.x, y0 = fma(t0, ob.w, ob.y) - p0.y;
                        const double x1 = fma(t1, ob.z, ob.x) - p1.x, y1 = fma(t1, ob.w, ob.y) - p1.y;
                        const double x2 = fma(t2, ob.z, ob.x) - p2.x, y2 = fma(t2, ob.w, ob.y) - p2.y;
                        const double x3 = fma(t3, ob.z, ob.x) - p3.x, y3 = fma(t3, ob.w, ob.y) - p3.y;
                        sweep_upd(fma(x0, x0, y0 * y0), j, b0, i0);
                        sweep_upd(fma(x1, x1, y1 * y1), j + 1, b1, i1);
                        sweep_upd(fma(x2, x2, y2 * y2), j + 2, b2, i2);
                        sweep_upd(fma(x3, x3, y3 * y3), j + 3, b3, i3);
                    }
                    for (; j < P; ++j, jd += 1.0) {
                        const double2 p0 = q[j];
                        const double t0 = jd;
                        const double x0 = fma(t0, ob.z, ob.x) - p0.x, y0 = fma(t0, ob.w, ob.y) - p0.y;
                        sweep_upd(fma(x0, x0, y0 * y0), j, b0, i0);
                    }
                    if (b1 < b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
                    if (b3 < b2 || (b3 == b2 && i3 < i2)) { b2 = b3; i2 = i3; }
                    if (b2 < b0 || (b2 == b0 && i2 < i0)) { b0 = b2; i0 = i2; }
                    const int bj = i0;
                    const double mx = fma((double)bj, ob.z, ob.x), my = fma((double)bj, ob.w, ob.y);
                    const int k = (bj == P - 1) ? P - 2 : bj;
                    double dd;
                    bestkey = min(bestkey, dp_owner_key(q[k], q[k + 1], bj, P, o, mx, my, lat_min, lat_max, &dd));
                }
            }
            const unsigned gmin = __reduce_min_sync(DP_FULL, bestkey);
            if (gmin != 0xffffffffu) {
                const int jstar = (int)(gmin >> 16);
                double t[(SWEEP_MAX_BASE + 31) / 32];
#pragma unroll
                for (int u = 0; u < (SWEEP_MAX_BASE + 31) / 32; ++u) {
                    const int j = lane + 32 * u;
                    t[u] = (j < jstar) ? sqrt(dp_sq2(q[j + 1].x - q[j].x, q[j + 1].y - q[j].y)) : 0.0;
                }
                __syncwarp();                               // every lane is done with q: reuse it for the terms
                double* len = reinterpret_cast<double*>(q);
#pragma unroll
                for (int u = 0; u < (SWEEP_MAX_BASE + 31) / 32; ++u) len[lane + 32 * u] = t[u];   // zero padded to 256
                __syncwarp();
                double sum = 0.0;
                for (int j = 0; j < jstar; j += 8) {        // index order; the zero padding does not change the sum
                    const double a0 = len[j], a1 = len[j + 1], a2 = len[j + 2], a3 = len[j + 3];
                    const double a4 = len[j + 4], a5 = len[j + 5], a6 = len[j + 6], a7 = len[j + 7];
                    sum += a0; sum += a1; sum += a2; sum += a3; sum += a4; sum += a5; sum += a6; sum += a7;
                }
                dis_lng = sum;
            }
        }
        if (lane == 0) {
            cand_dis_lng[c] = dis_lng;
            const float cost = (dis_lng > clear_dis) ? 0.0f : __int_as_float(0x7f800000);
            const unsigned long long key = ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)c;
            mykey = min(mykey, key);
        }
    }
    if (lane == 0) s_key[wib] = mykey;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = s_key[0];
        for (int w = 1; w < SWEEP_WARPS; ++w) k = min(k, s_key[w]);
        if (k != ~0ull) atomicMin(best_key, k);
    }
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
cudaError_t dp_launch_sweep(const double* base_x, const double* base_y, int n_base, const double* offset, const int32_t* n_pts,
                            int n_cand, const double* ox, const double* oy, const double* dvx, const double* dvy, int n_obs,
                            double lat_min, double lat_max, double clear_dis, double* cand_dis_lng, unsigned long long* best_key,
                            const int32_t* order, unsigned* next, cudaStream_t st) {
    if (n_cand <= 0) return cudaSuccess;
    int blocks = (n_cand + SWEEP_WARPS - 1) / SWEEP_WARPS;
    const int cap = 148 * SWEEP_CTAS_PER_SM;                               // persistent-style: 3 CTAs x 8 warps per SM, grid-stride over candidates
    if (blocks > cap) blocks = cap;
    sweep_kernel<<<blocks, SWEEP_WARPS * 32, 0, st>>>(base_x, base_y, n_base, offset, n_pts, n_cand, ox, oy, dvx, dvy, n_obs, lat_min,
                                                      lat_max, clear_dis, cand_dis_lng, best_key, order, next);
    return cudaGetLastError();
}
cudaError_t dp_launch_fma_peak(int which, float* sink, int iters, int blocks, cudaStream_t st) {
    if (which == 0) fma_peak_kernel<double><<<blocks, 256, 0, st>>>(reinterpret_cast<double*>(sink), iters);
    else fma_peak_kernel<float><<<blocks, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}
