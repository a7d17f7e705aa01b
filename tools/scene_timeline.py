"""debug: where does a one-wave cycle spend its time?  Needs a library built with `make DEBUG_CLOCK=1` (adds a
per-scene {start, end, smid} record per launch; never shipped).  Prints per-phase warp-duration percentiles, the
kernel span, and per-SM finish times of cycle K-1."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner, load  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 12
m = scenes.Map(); ep = scenes.Episodes(m, np.arange(n), cycles=K, n_obs=10); H, OX, OY = ep.all_cycles()
p = Planner(n, 10); p.upload_map(m)
if os.environ.get("TL_DEV"):                                  # the bench's `value` path: device pointers, L2 flushed before every cycle
    import torch
    dev = torch.device("cuda", 0)
    d_hdr = torch.from_numpy(H.view(np.uint8).reshape(K, n, 128)).to(dev); d_ox = torch.from_numpy(OX).to(dev); d_oy = torch.from_numpy(OY).to(dev)
    d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    for c in range(K):
        flush.fill_(c)
        p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
else:
    for c in range(K):
        o = p.cycle(np.ascontiguousarray(H[c]), OX[c], OY[c])
T2 = np.zeros((2, 65536, 2, 8), np.int64)
lib = load()
assert lib.dp_debug_scene_timeline(T2.ctypes.data_as(C.c_void_p), C.c_int(n)) == 0
last = int(np.argmax(T2[:, :n, 1, 1].max(axis=1)))          # parity of the last cycle
t = T2[last, :n]
t0 = t[:, 0, 0].min()
for ph, nm in ((0, "Decision"), (1, "Planning")):
    st, en, sm, nt = t[:, ph, 0] - t0, t[:, ph, 1] - t0, t[:, ph, 2], t[:, ph, 3]
    dur = (en - st) / 1e3
    print("%s: kernel span %.1f us (first start %.1f, last end %.1f); warp duration us p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f; start spread p99 %.1f us"
          % (nm, (en.max() - st.min()) / 1e3, st.min() / 1e3, en.max() / 1e3, *np.percentile(dur, [10, 50, 90, 99]), dur.max(),
             (np.percentile(st, 99) - st.min()) / 1e3))
    fin = np.array([en[sm == s].max() for s in np.unique(sm)]) / 1e3
    print("   per-SM finish us: min %.1f p50 %.1f p90 %.1f max %.1f ; SMs used %d ; warps/SM max %d"
          % (fin.min(), np.median(fin), np.percentile(fin, 90), fin.max(), np.unique(sm).size, np.bincount(sm).max()))
    for k in np.unique(nt):
        sel = nt == k
        if sel.sum() > n // 100:
            extra = ""
            if ph == 0:
                reg = (t[sel, 0, 4] - t[sel, 0, 0]) / 1e3
                extra = ", regions done at %.1f us" % reg.mean()
                sw = sel & (t[:, 0, 5] > 0)
                if sw.any():
                    extra += ", sweep %d scenes: start %.1f us, lasts %.1f us" % (sw.sum(), ((t[sw, 0, 5] - t[sw, 0, 0]) / 1e3).mean(),
                                                                              ((t[sw, 0, 6] - t[sw, 0, 5]) / 1e3).mean())
            print("   n_traj %2d: %5d scenes, mean duration %.1f us%s" % (k, sel.sum(), dur[sel].mean(), extra))
    busy = dur.sum() / (np.unique(sm).size * (en.max() - st.min()) / 1e3)
    print("   mean resident warps per SM over the span: %.1f" % busy)
