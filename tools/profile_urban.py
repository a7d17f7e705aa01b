"""Small driver for ncu / timelines: K cycles of the config-5 workload (urban junction, 200 agents, CTR tracks T = 400)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_obs = int(sys.argv[3]) if len(sys.argv) > 3 else 200
tracks = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
m = scenes.Map()
ep = scenes.Episodes(m, np.arange(n), cycles=K, n_obs=n_obs, kind="urban")
H, OX, OY, VX, VY, DTH = ep.all_cycles_tracks()
p = Planner(n, n_obs)
p.upload_map(m)
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
d_hdr = up(H.view(np.uint8).reshape(K, n, 128)); d_ox, d_oy, d_vx, d_vy, d_dth = up(OX), up(OY), up(VX), up(VY), up(DTH)
d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
for c in range(K):
    ev[c][0].record(st)
    if tracks:
        p.set_tracks_dev(ep.TRACK_T, d_vx[c].data_ptr(), d_vy[c].data_ptr(), d_dth[c].data_ptr())
    ev[c][1].record(st)
    p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
    ev[c][2].record(st)
torch.cuda.synchronize()
print("rollout ms:", ["%.3f" % e[0].elapsed_time(e[1]) for e in ev])
print("cycle ms:  ", ["%.3f" % e[1].elapsed_time(e[2]) for e in ev])
