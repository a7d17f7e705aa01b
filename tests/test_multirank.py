"""N > 1 host logic on CPU: two gloo ranks shard a scene batch by contiguous ranges, each runs its
shard (here with the CPU oracle standing in for the GPU), plan records are gathered, and the result is
byte-identical to the single-rank run.  Also the packed (cost, index) cross-rank argmin."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def worker(rank, world, port, n_total, cycles, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dmpp_b200  # noqa: F401
    from dmpp_b200 import parallel, scenes
    from oracle import binding
    m = scenes.Map()
    o = binding.Oracle(); o.set_map(m)
    lo, hi = parallel.scene_range(n_total, rank, world)
    ep = scenes.Episodes(m, np.arange(lo, hi), cycles=cycles)
    out = o.run(*ep.all_cycles(), paths=False, calls=False, trace=False, exhaustive=False)
    last = torch.from_numpy(out["rec"][-1].view(np.uint8).reshape(hi - lo, 128).copy())
    allrec = parallel.gather_plan_records(last)
    # split-candidate argmin: rank r proposes (cost, index); the lowest cost, then the lowest index, wins
    prop = [(0.0, 40), (0.0, 17)] if world == 2 else [(float("inf"), 3)] * world
    win = parallel.global_argmin(prop[rank][0], prop[rank][1])
    win_inf = parallel.global_argmin(float("inf"), 100 + rank)
    if rank == 0:
        q.put((allrec.numpy().tobytes(), win, win_inf))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])
def test_two_rank_shard_and_gather(n_total, oracle, the_map):
    from dmpp_b200 import scenes
    cycles = 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, n_total, cycles, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, win, win_inf = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ep = scenes.Episodes(the_map, np.arange(n_total), cycles=cycles)
    single = oracle.run(*ep.all_cycles(), paths=False, calls=False, trace=False, exhaustive=False)
    assert got == single["rec"][-1].tobytes()
    assert win == (0.0, 17)
    assert win_inf[0] == float("inf") and win_inf[1] == 100


def test_scene_range_partitions():
    from dmpp_b200.parallel import scene_range
    for n in (0, 1, 7, 4096, 1048576):
        for w in (1, 2, 4, 8):
            r = [scene_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
