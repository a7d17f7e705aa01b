"""debug: per-scene cycle counts by phase (library built with -DDP_DEBUG_CLOCK)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dmpp_b200
from dmpp_b200 import scenes, abi
from dmpp_b200.planner import Planner
n=int(sys.argv[1]) if len(sys.argv)>1 else 4096; K=12
m=scenes.Map(); ep=scenes.Episodes(m,np.arange(n),cycles=K,n_obs=10); H,OX,OY=ep.all_cycles()
p=Planner(n,10); p.upload_map(m)
for c in range(K):
    o=p.cycle(np.ascontiguousarray(H[c]),OX[c],OY[c])
r=o["rec"]; nt=r["n_traj"]
names=[("F search done","brakespeed"),("decision done","path_lat_dis"),("aim done","path_dir_err"),("bulk wait done","remain_dis"),("local state done","mindist_lat"),("path planned","mindist_lon"),("end","radius")]
for k in (3,5,11):
    sel=nt==k
    if sel.sum()==0: continue
    print("n_traj",k,"count",sel.sum(), " | ".join("%s %.0f"%(nm,r[f][sel].mean()) for nm,f in names))
print("all end mean %.0f max %.0f"%(r["radius"].mean(), r["radius"].max()))
