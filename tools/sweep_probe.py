"""config 3 latency floor: device / wall p50 of dp_sweep_score by rows (64, 1) and obstacle count (0, 1, 32, 50); with
DP_SWEEP_DBG=1 also the in-kernel phase stamps of the row owners (dp_sweep_debug)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmpp_b200
from dmpp_b200 import scenes
from dmpp_b200.planner import Planner
m = scenes.Map()
rng = np.random.default_rng(2024)
gl = m.lane_index(3, 2); o = m.lane_pt_off[gl] + 900
bx, by = m.x[o:o + 256], m.y[o:o + 256]
def grid(nlat):
    lat = -3.15 + 0.1 * np.arange(nlat); aim = 10.0 + 2.5 * np.arange(32); hor = 8 * (1 + np.arange(32))
    L, A, Hh = np.meshgrid(lat, aim, hor, indexing="ij")
    return L.ravel(), np.minimum(Hh, np.maximum(2, (A / 0.5).astype(np.int64))).astype(np.int32).ravel()
p = Planner(16, 64); p.upload_map(m)
for nlat in (64, 1):
    offset, n_pts = grid(nlat)
    sess = p.sweep_session(bx, by, offset, n_pts, 64)
    for N in (0, 1, 32, 50):
        idx = rng.integers(5, 250, N)
        ox0, oy0 = bx[idx] + rng.normal(0, 1.5, N), by[idx] + rng.normal(0, 1.5, N)
        dvx, dvy = rng.normal(0, 0.04, N), rng.normal(0, 0.04, N)
        for i in range(50): sess.score(ox0, oy0, dvx, dvy)
        dev = np.array([sess.score(ox0, oy0, dvx, dvy, want_dis=False)[2] for i in range(1000)])
        w = np.zeros(1000)
        for i in range(1000):
            t0 = time.perf_counter(); sess.score(ox0, oy0, dvx, dvy, want_ms=False); w[i] = time.perf_counter() - t0
        print("rows %d n_obs %d: device p50 %.1f us, wall p50 %.1f us" % (nlat, N, np.median(dev) * 1e3, np.median(w) * 1e6), flush=True)
    sess.close()
if os.environ.get("DP_SWEEP_DBG"):
    import ctypes as C
    offset, n_pts = grid(64)
    sess = p.sweep_session(bx, by, offset, n_pts, 64)
    N = 50; idx = rng.integers(5, 250, N)
    ox0, oy0 = bx[idx] + rng.normal(0, 1.5, N), by[idx] + rng.normal(0, 1.5, N)
    dvx, dvy = rng.normal(0, 0.04, N), rng.normal(0, 0.04, N)
    acc = []
    for i in range(200):
        sess.score(ox0, oy0, dvx, dvy)
        T = np.zeros((64, 8), np.int64)
        assert p.lib.dp_sweep_debug(sess.h, T.ctypes.data_as(C.c_void_p), C.c_int(64)) == 0
        t0 = T[:, 0].min()
        last = int(np.argmax(T[:, 4]))
        acc.append([T[:, 0].max() - t0, np.median(T[:, 1]) - t0, T[:, 1].max() - t0, T[:, 3].max() - t0, T[:, 4].max() - t0, T[last, 5] - t0, T[last, 6] - t0])
    a = np.median(np.array(acc[20:]), axis=0) / 1e3
    print("us since the first CTA started: last CTA start %.2f | row pass done p50 %.2f max %.2f | select+reduce done max %.2f | counted max %.2f | "
          "winner known %.2f | system fence done %.2f" % tuple(a))
