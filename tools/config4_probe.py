"""config 4 shard through bench.config4_line under torchrun, gather mode from DP_CFG4_GATHER = deferred | eager | none"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Planner  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
r = bench.config4_line(torch, dist, Planner, scenes.Map(), dev, rank, world, lr, total)
if rank == 0:
    print(json.dumps({k: r[k] for k in ("scenes_per_gpu", "value", "ms_per_step", "gather")}))
dist.destroy_process_group()
