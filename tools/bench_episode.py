"""whole 25-cycle episodes resident in HBM through dp_run_episode_dev (one call) against 25 separate
dp_cycle_batch_dev calls; no L2 flush, CUDA events around the whole episode.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cycles, reps = 25, 8
m = scenes.Map(); ep = scenes.Episodes(m, np.arange(n), cycles=cycles, n_obs=10); H, OX, OY = ep.all_cycles()
p = Planner(n, 10); p.upload_map(m)
dev = torch.device("cuda", 0)
d_h = torch.from_numpy(H.view(np.uint8).reshape(cycles, n, 128)).to(dev); d_x = torch.from_numpy(OX).to(dev); d_y = torch.from_numpy(OY).to(dev)
d_r = torch.zeros((cycles, n, 128), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
for mode in ("episode", "per_cycle"):
    ms = []
    for r in range(reps):
        p.reset(0, n); torch.cuda.synchronize()
        e0.record(st)
        if mode == "episode":
            p.run_episode_dev(n, cycles, d_h.data_ptr(), d_x.data_ptr(), d_y.data_ptr(), d_r.data_ptr(), stream=st.cuda_stream)
        else:
            for c in range(cycles):
                p.cycle_dev(n, d_h[c].data_ptr(), d_x[c].data_ptr(), d_y[c].data_ptr(), d_r[c].data_ptr(), stream=st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = float(np.median(ms[2:]))
    out[mode] = {"ms_per_episode": best, "us_per_cycle": best / cycles * 1e3, "plan_cycles_per_s": n * cycles / (best * 1e-3)}
print(json.dumps({"scenes": n, "cycles": cycles, **out}))
