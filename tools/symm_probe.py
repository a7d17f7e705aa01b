"""probe: does torch symmetric memory (peer-mapped buffers + barrier) work on this box?  torchrun, 2+ ranks."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty((world * 1024,), dtype=torch.int32, device=dev)
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", hex(h.multicast_ptr) if h.has_multicast_support else None, flush=True)
h.barrier()
for p in range(world):                                   # write my slice into every peer's buffer
    peer = h.get_buffer(p, (world * 1024,), torch.int32)
    peer[rank * 1024:(rank + 1) * 1024] = rank + 1
h.barrier()
torch.cuda.synchronize()
print(rank, "gathered ok:", bool((t.view(world, 1024) == torch.arange(1, world + 1, device=dev, dtype=torch.int32)[:, None]).all()), flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5): h.barrier()
ev0.record()
for _ in range(100): h.barrier()
ev1.record(); torch.cuda.synchronize()
print(rank, "barrier us", ev0.elapsed_time(ev1) * 10, flush=True)
dist.destroy_process_group()
