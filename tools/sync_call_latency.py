import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import dmpp_b200
from dmpp_b200 import abi, scenes
from dmpp_b200.planner import Planner
n=4096; K=25
m=scenes.Map(); ep=scenes.Episodes(m,np.arange(n),cycles=K,n_obs=10); H,OX,OY=ep.all_cycles()
p=Planner(n,10); p.upload_map(m)
pin=lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
Hh=pin(H.view(np.uint8).reshape(K,n,128)).view(abi.scene_hdr).reshape(K,n); OXh,OYh=pin(OX),pin(OY)
rec=torch.empty((n,128),dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(n)
for mode in ("cycle","submit+wait"):
    ts=[]
    for rep in range(3):
        p.reset(0,n)
        for c in range(K):
            t0=time.perf_counter()
            if mode=="cycle": p.cycle(Hh[c],OXh[c],OYh[c],out={"rec":rec})
            else: p.submit(Hh[c],OXh[c],OYh[c],rec); p.wait()
            ts.append(time.perf_counter()-t0)
    ts=np.array(ts[K:])*1e6
    print(mode, "us per call: mean %.1f p50 %.1f min %.1f"%(ts.mean(),np.median(ts),ts.min()))
