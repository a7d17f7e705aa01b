"""What the fused gather adds to a 4096-scene cycle at N > 1, leg by leg (run under torchrun, one rank per GPU):
  plain    no mirrors, no flags                        (the N = 1 kernel)
  local    records mirrored into this rank's OWN gathered buffer + flags (no NVLink traffic)
  peers    records + flags stored on every rank        (dp_gather_arm)
  chained  ... and the wait for the previous step folded into the launch (dp_gather_chain)
  deferred the launch forwards the records of the launch BEFORE as its warps start, raises that step's flags when its Decision
           half retires and waits for them in its last warp (dp_gather_arm_deferred)
Each leg: 200 cycles, 256 MiB L2 flush before each, CUDA events around the launch; prints p50 / mean per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from dmpp_b200.planner import Gather, Planner  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n, n_obs, ep_len, reps = 4096, 10, 25, 200
    m = scenes.Map()
    ep = scenes.Episodes(m, np.arange(rank * n, (rank + 1) * n), cycles=ep_len, n_obs=n_obs)
    H, OX, OY = ep.all_cycles()
    p = Planner(n, n_obs, device=lr)
    p.upload_map(m)
    g = Gather(p, world, rank, n, depth=4)
    hs = [None] * world
    dist.all_gather_object(hs, g.my_handle())
    g.attach(hs)
    d_hdr = torch.from_numpy(H.view(np.uint8).reshape(ep_len, n, 128)).to(dev)
    d_ox, d_oy = torch.from_numpy(OX).to(dev), torch.from_numpy(OY).to(dev)
    d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    step = [0]

    def leg(name):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize(); dist.barrier()
        for i in range(reps):
            c = i % ep_len
            if c == 0:
                p.reset_dev(0, n, stream=st.cuda_stream)     # (ordered with the launches of this loop)
            flush.zero_()
            if name == "plain":
                g.disarm()
            elif name == "local":
                g.disarm()
                p.set_record_mirrors([g.buffer(1)])
            elif name == "deferred":
                step[0] += 1
                g.arm_deferred(step[0])
            else:
                step[0] += 1
                g.arm(step[0])
                g.chain(step[0] - 1 if (name == "chained" and i) else 0)
            ev[i][0].record(st)
            p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
            ev[i][1].record(st)
        if name in ("peers", "chained"):
            g.wait(step[0], stream=st.cuda_stream)
        if name == "deferred":
            g.flush(stream=st.cuda_stream)
        torch.cuda.synchronize(); dist.barrier()
        p.set_record_mirrors([])
        ms = np.array([a.elapsed_time(b) for a, b in ev])[5:]
        print("rank %d %-8s p50 %.2f us  mean %.2f us  p90 %.2f us" % (rank, name, np.percentile(ms, 50) * 1e3, ms.mean() * 1e3, np.percentile(ms, 90) * 1e3), flush=True)

    for name in os.environ.get("LEGS", "plain,local,peers,chained,deferred,plain").split(","):
        leg(name)
    g.close(); p.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
