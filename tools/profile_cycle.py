"""Small driver for ncu: W warm-up cycles then K cycles of the config-2 workload (inputs resident in HBM)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
m = scenes.Map()
ep = scenes.Episodes(m, np.arange(n), cycles=K, n_obs=10)
H, OX, OY = ep.all_cycles()
p = Planner(n, 10)
p.upload_map(m)
dev = torch.device("cuda", 0)
d_hdr = torch.from_numpy(H.view(np.uint8).reshape(K, n, 128)).to(dev)
d_ox = torch.from_numpy(OX).to(dev)
d_oy = torch.from_numpy(OY).to(dev)
d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for c in range(K):
    ev[c][0].record(st)
    p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
    ev[c][1].record(st)
torch.cuda.synchronize()
print("kernel ms per cycle:", ["%.3f" % e[0].elapsed_time(e[1]) for e in ev])
if len(sys.argv) > 3:
    from dmpp_b200 import abi
    p.reset(0, n)
    for c in range(K):
        p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
        r = d_rec.cpu().numpy().view(abi.plan_record).reshape(n)
        nt = r["n_traj"]
        print("cycle", c, "traj", int(nt.sum()), "hist", dict(zip(*[x.tolist() for x in np.unique(nt, return_counts=True)])),
              "afresh", int(r["afresh_planning"].sum()), "dlg", dict(zip(*[x.tolist() for x in np.unique(r["behavior_to_dlg"], return_counts=True)])))
