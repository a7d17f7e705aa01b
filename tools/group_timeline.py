"""Phase timeline of the group kernel (csrc/dp_group.cuh): runs K cycles of the config-2 workload with DP_TIMELINE=1 and
prints, per phase, the mean / max time (over CTAs) between consecutive phase stamps of the LAST cycle, plus the CTA spans.
usage: python tools/group_timeline.py [scenes] [cycles]"""
import ctypes as C
import os
import sys

os.environ["DP_TIMELINE"] = "1"
os.environ.setdefault("DP_KERNEL", "group")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import abi, scenes  # noqa: E402
from dmpp_b200.planner import Planner, _ck  # noqa: E402

NAMES = {1: "P0 load hdr/carry, tables", 2: "P1 preamble (scene)", 3: "P2 region scans (2 passes)", 4: "P3 region terms + sums",
         5: "P4 rule tree (scene)", 6: "P5 sweep scans", 7: "P5b sweep terms + sums", 8: "P6 decision out + aim recipe (scene)",
         9: "P7 walk via prefix table + nearest partials", 10: "P8 aim point (scene)", 11: "P8b first Bezier", 12: "P8c first nearest",
         13: "P9 local state (scene)", 14: "P10 remain terms", 15: "P11 remain sum, judge (scene)", 16: "P11b mean terms",
         17: "P11c mean prefix", 18: "P12 Bezier / MeanPoints points", 19: "P13 local bounds + path out", 20: "P14 local scan (2 passes)",
         21: "P14b local terms", 22: "P15 local sum, speed (scene)", 23: "P16 stores"}

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
for c in range(K):
    p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
nb = min(8192, n)
tl = np.zeros((nb, 32), np.int64)
_ck(p.lib.dp_debug_timeline(p.ctx, abi.ptr(tl), C.c_int(nb)), "dp_debug_timeline")
tl = tl[tl[:, 0] > 0]
print("CTAs recorded:", tl.shape[0])
t0 = tl[:, 0].min()
end = tl[:, 23]
print("kernel span %.1f us; CTA start spread %.1f us; CTA duration mean %.1f max %.1f us" % (
    (end.max() - t0) / 1e3, (tl[:, 0].max() - t0) / 1e3, (end - tl[:, 0]).mean() / 1e3, (end - tl[:, 0]).max() / 1e3))
prev = tl[:, 0]
for i in range(1, 24):
    cur = tl[:, i]
    ok = cur > 0
    if not ok.any():
        continue
    d = (cur - prev)[ok] / 1e3
    print("%-40s ctas %4d  mean %6.2f  p90 %6.2f  max %6.2f us" % (NAMES[i], ok.sum(), d.mean(), np.percentile(d, 90), d.max()))
    prev = np.where(ok, cur, prev)
