"""Device time per cycle of one workload shape under the library / kernel given by the environment (DP_DEBUG_LIB, DP_KERNEL,
DP_GROUP_CFG): python tools/compare_kernels.py <scenes> <n_obs> <kind> [cycles]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner  # noqa: E402

n, n_obs, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
K = int(sys.argv[4]) if len(sys.argv) > 4 else 12
m = scenes.Map()
ep = scenes.Episodes(m, np.arange(n), cycles=K, n_obs=n_obs, kind=kind)
H, OX, OY = ep.all_cycles()
p = Planner(n, n_obs)
p.upload_map(m)
dev = torch.device("cuda", 0)
d_hdr = torch.from_numpy(H.view(np.uint8).reshape(K, n, 128)).to(dev)
d_ox = torch.from_numpy(OX).to(dev)
d_oy = torch.from_numpy(OY).to(dev)
d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()
ms = []
for rep in range(2):
    p.reset(0, n)
    for c in range(K):
        flush.fill_(c)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        if rep == 1 and c >= 3:
            ms.append(e0.elapsed_time(e1))
print("%s n=%d n_obs=%d kind=%s: median %.1f us  min %.1f  max %.1f (cold L2, cycles 3..%d)" % (
    os.environ.get("DP_DEBUG_LIB", "product").split("/")[-1] + ":" + os.environ.get("DP_KERNEL", "default"), n, n_obs, kind,
    1e3 * float(np.median(ms)), 1e3 * min(ms), 1e3 * max(ms), K - 1))
