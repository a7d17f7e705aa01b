// csrc/dp_world.cu -- the two stages either side of the Decision + Planning cycle:
//   dp_world_kernel   the world step of the closed-loop episode runner: ego walked along its own plan, agents along their
//                     lanes, windowed re-localisation; one warp per scene, everything in place in HBM
//                     (arithmetic frozen in oracle/world_spec.cpp; IEEE binary64, -fmad=false, operations in that order);
//   dp_frames_kernel  the output stage: PlanningOut / PlanningStatus (Planning.cpp:173-214) packed into fixed-layout frames
//                     from the plan record and the carried local path -- byte work bound by HBM: a warp streams one scene
//                     (3.3 KB in, 3.3 KB out), lanes own consecutive 16-byte points so that loads and stores coalesce.
#include "dp_kernels.h"

namespace {
constexpr int WPB = 4;   // warps (= scenes) per CTA

__device__ __forceinline__ double seg_len(const double2 a, const double2 b, double* ddx, double* ddy) {
    *ddx = b.x - a.x; *ddy = b.y - a.y;
    return sqrt(*ddx * *ddx + *ddy * *ddy);
}

__global__ void __launch_bounds__(WPB * 32) dp_world_kernel(DevMap m, dp_params p, dp_world_params wp, int n, int max_obs,
                                                            dp_scene_hdr* __restrict__ hdr, dp_agent* __restrict__ agents,
                                                            double* __restrict__ obs_x, double* __restrict__ obs_y,
                                                            const dp_plan_record* __restrict__ rec, const double2* __restrict__ last_path) {
    __shared__ __align__(16) uint32_t sh[WPB][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int scene = blockIdx.x * WPB + wib;
    if (scene >= n) return;
    sh[wib][lane] = reinterpret_cast<const uint32_t*>(hdr + scene)[lane];   // one coalesced 128-byte load
    __syncwarp();
    const dp_scene_hdr& h = *reinterpret_cast<const dp_scene_hdr*>(sh[wib]);
    const int n_obs = h.n_obs;
    const double dt = h.period_ms / 1000.0;
    double x = h.x, y = h.y, dir = h.dir, vn = h.velocity;
    const bool step = rec != nullptr;
    if (step && lane == 0) {
        // ---- ego: bounded approach to the planned speed, then a walk along the carried local path ----
        const dp_plan_record& r = rec[scene];
        const int gl = m.road_lane_base[h.road_num - 1] + h.lane_num - 1;
        const int cnt = m.lane_pt_off[gl + 1] - m.lane_pt_off[gl];
        const bool at_end = h.id[h.lane_num - 1] >= cnt - wp.end_margin;
        const double v = h.velocity, vt = r.brakespeed * 3.6, dvm = wp.a_max * 3.6 * dt;
        double dv = vt - v;
        if (dv > dvm) dv = dvm;
        if (dv < -dvm) dv = -dvm;
        vn = v + dv;
        if (vn < 0) vn = 0;
        if (at_end) vn = 0;
        const double ds = at_end ? 0.0 : (v + vn) * 0.5 / 3.6 * dt;
        if (ds > 0) {
            const double2* lp = last_path + (size_t)scene * DP_PATH_POINTS;
            int i = r.afresh_planning ? 0 : r.path_near_id;
            if (i < 0) i = 0;
            if (i > DP_PATH_POINTS - 2) i = DP_PATH_POINTS - 2;
            double rem = ds;
            double2 a = lp[i];
            for (;;) {
                const double2 b = lp[i + 1];
                double ddx, ddy;
                const double L = seg_len(a, b, &ddx, &ddy);
                if (rem < L || i == DP_PATH_POINTS - 2) {
                    if (L > 0) {
                        double t = rem / L;
                        if (t > 1) t = 1;
                        x = a.x + t * ddx; y = a.y + t * ddy;
                        dir = dp_heading(a.x, a.y, b.x, b.y, p.epsilon, p.pi);
                    } else { x = a.x; y = a.y; }
                    break;
                }
                rem -= L; ++i; a = b;
            }
        }
    }
    x = __shfl_sync(DP_FULL, x, 0); y = __shfl_sync(DP_FULL, y, 0);
    // ---- agents: constant speed along their lanes, then their obstacle points ----
    for (int k = lane; k < n_obs; k += 32) {
        dp_agent a = agents[(size_t)scene * max_obs + k];
        const int off = m.lane_pt_off[a.lane], cntk = m.lane_pt_off[a.lane + 1] - off;
        double ddx, ddy;
        if (step) {
            double u = a.u + a.v * dt;
            int i = a.i;
            while (i < cntk - 2) {
                const double L = seg_len(m.xy[off + i], m.xy[off + i + 1], &ddx, &ddy);
                if (u < L) break;
                u -= L; ++i;
            }
            if (i >= cntk - 2) {
                i = cntk - 2;
                const double L = seg_len(m.xy[off + i], m.xy[off + i + 1], &ddx, &ddy);
                if (u > L) u = L;
            }
            a.u = u; a.i = i;
            agents[(size_t)scene * max_obs + k] = a;
        }
        const double2 q = m.xy[off + a.i];
        const double L = seg_len(q, m.xy[off + a.i + 1], &ddx, &ddy);
        double ox = q.x, oy = q.y;
        if (L > 0) {
            const double t = a.u / L;
            const double px = q.x + t * ddx, py = q.y + t * ddy;
            ox = px + a.lat * (-(ddy / L));
            oy = py + a.lat * (ddx / L);
        }
        obs_x[(size_t)scene * max_obs + k] = ox;
        obs_y[(size_t)scene * max_obs + k] = oy;
    }
    // ---- localisation: nearest point per lane inside a window around the previous index, nearest lane ----
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int base = m.road_lane_base[h.road_num - 1];
    int nl = m.road_lane_base[h.road_num] - base;
    if (nl > DP_LANESUM) nl = DP_LANESUM;
    double best = INF;
    int bl = 0, my_id = 0;                                  // lane l keeps the new id of map lane l
    for (int l = 0; l < nl; ++l) {
        const int off = m.lane_pt_off[base + l], cnt = m.lane_pt_off[base + l + 1] - off;
        int lo = h.id[l] - wp.loc_back, hi = h.id[l] + wp.loc_fwd;
        if (lo < 0) lo = 0;
        if (hi > cnt - 1) hi = cnt - 1;
        double be = INF;
        int bi = 0x7fffffff;
        for (int i = lo + lane; i <= hi; i += 32) {
            const double2 q = m.xy[off + i];
            const double dx = x - q.x, dy = y - q.y;
            const double e = dx * dx + dy * dy;
            if (e < be) { be = e; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oe = __shfl_xor_sync(DP_FULL, be, o);
            const int oi = __shfl_xor_sync(DP_FULL, bi, o);
            if (oe < be || (oe == be && oi < bi)) { be = oe; bi = oi; }
        }
        if (bi == 0x7fffffff) bi = lo;
        if (lane == l) my_id = bi;
        if (be < best) { best = be; bl = l; }
    }
    // ---- the header of the next cycle, in place ----
    dp_scene_hdr* out = hdr + scene;
    if (lane < nl) out->id[lane] = my_id;
    if (lane == 0) {
        out->x = x; out->y = y; out->dir = dir; out->velocity = vn;
        out->lane_num = (uint16_t)(bl + 1);
    }
}

__global__ void __launch_bounds__(WPB * 32) dp_frames_kernel(dp_params p, int n, const dp_plan_record* __restrict__ rec,
                                                             const double2* __restrict__ last_path, dp_ctrl_frame* __restrict__ ctrl,
                                                             dp_status_frame* __restrict__ status) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int scene = blockIdx.x * WPB + wib;
    if (scene >= n) return;
    const dp_plan_record* r = rec + scene;
    // heads: lane 0 (64-bit stores into zeroed-by-construction layouts: every byte of the head is written here)
    if (lane == 0) {
        const double lon = r->mindist_lon, bs = r->brakespeed, da = r->des_acc;
        const unsigned light = r->light;
        if (ctrl) {
            dp_ctrl_frame* c = ctrl + scene;
            const unsigned long long w0 = (unsigned long long)r->cnt | ((unsigned long long)r->acc_flag << 16) | (1ull << 40) |
                                          ((unsigned long long)light << 48);          // cnt, apa 0, desacc_vd, desstr_vd 0, road_type 0, sstop 1, light
            *reinterpret_cast<unsigned long long*>(c) = w0;
            c->brakedis = lon; c->brake_speed = 0.0; c->desacc = da; c->desspd = bs; c->desstr = 0.0; c->radius = r->radius;
        }
        if (status) {
            dp_status_frame* s = status + scene;
            const unsigned long long w0 = (unsigned long long)(unsigned)(int)r->afresh_cause | ((unsigned long long)light << 32);
            *reinterpret_cast<unsigned long long*>(s) = w0;
            s->near_ob_dist = lon; s->planspeed = bs; s->planacc = da;
        }
    }
    // every 2nd path point (Planning.cpp:180-183, :203-212); frames are 8-byte aligned only, so points go out as two doubles
    const double2* lp = last_path + (size_t)scene * DP_PATH_POINTS;
    for (int i = lane; i < DP_OUT_POINTS; i += 32) {
        const double2 q = lp[2 * i];
        if (ctrl) {
            double* o = &ctrl[scene].pnts[i][0];
            o[0] = fma(q.y, p.k_lat, p.lat0);               // GlobalToWGS84 (operator specification): lat, then lng
            o[1] = fma(q.x, p.k_lng, p.lng0);
        }
        if (status) {
            double* o = &status[scene].path_points[i][0];
            o[0] = q.x; o[1] = q.y;
        }
    }
}
}  // namespace

cudaError_t dp_launch_world(const DevMap& m, const dp_params& p, const dp_world_params& wp, int n, int max_obs, dp_scene_hdr* hdr,
                            dp_agent* agents, double* obs_x, double* obs_y, const dp_plan_record* rec, const double2* last_path,
                            cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    dp_world_kernel<<<(n + WPB - 1) / WPB, WPB * 32, 0, st>>>(m, p, wp, n, max_obs, hdr, agents, obs_x, obs_y, rec, last_path);
    return cudaGetLastError();
}
cudaError_t dp_launch_frames(const dp_params& p, int n, const dp_plan_record* rec, const double2* last_path, dp_ctrl_frame* ctrl,
                             dp_status_frame* status, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    dp_frames_kernel<<<(n + WPB - 1) / WPB, WPB * 32, 0, st>>>(p, n, rec, last_path, ctrl, status);
    return cudaGetLastError();
}
