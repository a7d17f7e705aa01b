"""debug (library built with `make debug`, DP_DEBUG_LIB): run the bench workload up to cycle K of an episode and list the
slowest scenes of that cycle (Decision / Planning spans, n_traj, position in the kernel timeline)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner, load  # noqa: E402

n = 4096
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = scenes.Map(); ep = scenes.Episodes(m, np.arange(n), cycles=25, n_obs=10); H, OX, OY = ep.all_cycles()
p = Planner(n, 10); p.upload_map(m)
lib = load()
for c in range(K + 1):
    o = p.cycle(np.ascontiguousarray(H[c]), OX[c], OY[c])
T2 = np.zeros((2, 65536, 2, 8), np.int64)
assert lib.dp_debug_scene_timeline(T2.ctypes.data_as(C.c_void_p), C.c_int(n)) == 0
last = int(np.argmax(T2[:, :n, 1, 1].max(axis=1)))
t = T2[last, :n]
t0 = t[:, 0, 0].min()
dec = (t[:, 0, 1] - t[:, 0, 0]) / 1e3; pl = (t[:, 1, 1] - t[:, 1, 0]) / 1e3; end = (t[:, 1, 1] - t0) / 1e3
print("cycle %d: kernel pair ends at %.1f us; Decision max %.1f us, Planning max %.1f us" % (K, end.max(), dec.max(), pl.max()))
r = o["rec"]
for s in np.argsort(-end)[:8]:
    h = H[K][s]
    print("scene %4d: Decision %.1f us (start %.1f) Planning %.1f us (start %.1f, end %.1f) n_traj %d behavior %d afresh %d cause %d road %d lane %d id %s pos %d"
          % (s, dec[s], (t[s, 0, 0] - t0) / 1e3, pl[s], (t[s, 1, 0] - t0) / 1e3, end[s], r["n_traj"][s], r["behavior"][s], r["afresh_planning"][s],
             r["afresh_cause"][s], h["road_num"], h["lane_num"], h["id"][:3].tolist(), h["pos"]))
