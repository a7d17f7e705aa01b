"""Multi-GPU plumbing: scenes are independent, so the compute phase has no collective.
torch.distributed is used only to (a) gather the fixed-size plan records of all ranks and (b) pick the
global argmin when ONE scene's candidate set is split across ranks.  Backend: nccl on GPUs, gloo in the
CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def scene_range(n_total, rank, world):
    """contiguous scene-index range [lo, hi) of `rank` (same split the oracle's thread pool uses)"""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def gather_plan_records(rec_local, group=None):
    """rec_local: uint8 tensor [n_local, 128] (dp_plan_record rows) -> [sum n_local, 128] on every rank.
    Ranks may hold different counts (ragged last shard): counts are exchanged first."""
    world = dist.get_world_size(group)
    n = torch.tensor([rec_local.shape[0]], dtype=torch.int64, device=rec_local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0], rec_local.shape[1]), dtype=rec_local.dtype, device=rec_local.device)
        dist.all_gather_into_tensor(out, rec_local.contiguous(), group=group)
        return out
    mx = max(counts)
    pad = torch.zeros((mx, rec_local.shape[1]), dtype=rec_local.dtype, device=rec_local.device)
    pad[: rec_local.shape[0]] = rec_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def pack_cost_index(cost, index):
    """(float32 cost bits << 32) | index as int64: for cost >= 0 the IEEE order equals the integer order, so an
    integer MIN is an argmin with a deterministic lowest-index tie-break; infeasible = +inf."""
    bits = np.asarray(cost, dtype=np.float32).view(np.uint32).astype(np.int64)
    return (bits << 32) | np.asarray(index, dtype=np.int64)


def global_argmin(cost, index, device="cpu", group=None):
    """cost >= 0 (float), index: this rank's best candidate (global index).  Returns (cost, index) of the winner."""
    key = torch.tensor([int(pack_cost_index(cost, index))], dtype=torch.int64, device=device)
    dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    k = int(key.item())
    c = np.array([k >> 32], dtype=np.uint32).view(np.float32)[0]
    return float(c), int(k & 0xFFFFFFFF)
