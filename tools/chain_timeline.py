"""debug: how do consecutive cycles overlap in the chained submit path?  Needs `make debug` (instrumented library).
Runs K pipelined cycles of n scenes and prints, for the last two cycles, when their Decision / Planning warps started and
ended (GPU global timer, microseconds from the start of the older cycle's Decision launch)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import abi, scenes  # noqa: E402
from dmpp_b200.planner import Planner, load  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 14
m = scenes.Map(); ep = scenes.Episodes(m, np.arange(n), cycles=K, n_obs=10); H, OX, OY = ep.all_cycles()
p = Planner(n, 10); p.upload_map(m)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
Hh = pin(H.view(np.uint8).reshape(K, n, 128)).view(abi.scene_hdr).reshape(K, n); OXh, OYh = pin(OX), pin(OY)
recs = [torch.empty((n, 128), dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(n) for _ in range(2)]
p.submit(Hh[0], OXh[0], OYh[0], recs[0])
for c in range(1, K):
    p.submit(Hh[c], OXh[c], OYh[c], recs[c & 1]); p.wait()
p.wait()
T = np.zeros((2, 65536, 2, 8), np.int64)
assert load().dp_debug_timeline(T.ctypes.data_as(C.c_void_p), C.c_int(n)) == 0
T = T[:, :n]
new = int(np.argmax(T[:, :, 1, 1].max(axis=1))); old = 1 - new
t0 = T[old, :, 0, 0].min()
for nm, par in (("older cycle", old), ("newer cycle", new)):
    for ph, pn in ((0, "Decision"), (1, "Planning")):
        st = (T[par, :, ph, 0] - t0) / 1e3; en = (T[par, :, ph, 1] - t0) / 1e3
        print("%s %s: start min %.1f p50 %.1f p99 %.1f | end p50 %.1f p99 %.1f max %.1f" % (nm, pn, st.min(), np.median(st), np.percentile(st, 99),
                                                                                       np.median(en), np.percentile(en, 99), en.max()))
print("period (newer Planning end - older Planning end): %.1f us" % ((T[new, :, 1, 1].max() - T[old, :, 1, 1].max()) / 1e3))
