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
recs=[torch.empty((n,128),dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(n) for _ in range(2)]
ts=[];tw=[];tsum=[]
for rep in range(6):
    p.reset(0,n)
    p.submit(Hh[0],OXh[0],OYh[0],recs[0])
    t_all=time.perf_counter()
    for c in range(1,K):
        t0=time.perf_counter(); p.submit(Hh[c],OXh[c],OYh[c],recs[c&1]); t1=time.perf_counter(); p.wait(); t2=time.perf_counter()
        s=int(recs[(c-1)&1]["n_traj"].sum(dtype=np.int64)) if not os.environ.get("NOSUM") else 0; t3=time.perf_counter()
        if rep>=2: ts.append(t1-t0); tw.append(t2-t1); tsum.append(t3-t2)
    p.wait()
    if rep>=2: print("per step us", (time.perf_counter()-t_all)/(K-1)*1e6)
print("submit us mean %.1f p50 %.1f | wait us mean %.1f p50 %.1f | sum us %.1f"%(np.mean(ts)*1e6,np.median(ts)*1e6,np.mean(tw)*1e6,np.median(tw)*1e6,np.mean(tsum)*1e6))
