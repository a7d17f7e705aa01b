"""Fused record gather through the C ABI alone (dp_gather_*: CUDA IPC peer mapping + completion flags raised by the cycle kernel).
Two processes share cuda:0 as "ranks" of a two-GPU job: after each step both must hold both ranks' plan records, bit for bit, for
both kernels (N = 10: warp-per-scene, N = 40: group)."""
import multiprocessing as mp

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def worker(rank, world, conn, n, n_obs, cycles, q):
    import ctypes as C
    import torch
    import dmpp_b200  # noqa: F401
    from dmpp_b200 import abi, scenes
    from dmpp_b200.planner import Gather, Planner
    dev = torch.device("cuda", 0)
    m = scenes.Map()
    ep = scenes.Episodes(m, np.arange(rank * n, (rank + 1) * n), cycles=cycles, n_obs=n_obs)
    H, OX, OY = ep.all_cycles()
    p = Planner(n, n_obs)
    p.upload_map(m)
    g = Gather(p, world, rank, n, depth=4)
    conn.send(g.my_handle())
    handles = conn.recv()                                    # all ranks' handles, index = rank
    g.attach(handles)
    d_hdr = torch.from_numpy(H.view(np.uint8).reshape(cycles, n, 128)).to(dev)
    d_ox, d_oy = torch.from_numpy(OX).to(dev), torch.from_numpy(OY).to(dev)
    d_rec = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    got = []
    for c in range(cycles):
        step = c + 1
        g.arm(step)
        p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
        g.wait(step, stream=st.cuda_stream)
        host = np.zeros((world * n, 128), np.uint8)
        assert p.lib.dp_memcpy_d2h(p.ctx, abi.ptr(host), C.c_void_p(g.buffer(step)), C.c_size_t(host.nbytes), C.c_void_p(st.cuda_stream)) == 0
        assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
        got.append((host.copy(), d_rec.cpu().numpy().copy()))
    # the host-pointer call (pinned buffers -> dp_cycle_submit + dp_cycle_wait under the hood) raises the flags too
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    Hh = pin(H.view(np.uint8).reshape(cycles, n, 128)).view(abi.scene_hdr).reshape(cycles, n)
    OXh, OYh = pin(OX), pin(OY)
    rec_h = torch.empty((n, 128), dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(n)
    p.reset(0, n)
    for c in range(cycles):
        step = cycles + c + 1
        g.arm(step)
        p.cycle(Hh[c], OXh[c], OYh[c], out={"rec": rec_h})
        g.wait(step, stream=st.cuda_stream)
        host = np.zeros((world * n, 128), np.uint8)
        assert p.lib.dp_memcpy_d2h(p.ctx, abi.ptr(host), C.c_void_p(g.buffer(step)), C.c_size_t(host.nbytes), C.c_void_p(st.cuda_stream)) == 0
        assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
        got.append((host.copy(), rec_h.view(np.uint8).reshape(n, 128).copy()))
    # chained: no wait launches -- the launch of step s carries the wait for step s-1 (dp_gather_chain); the gathered buffer of
    # s-1 is read right behind the kernel of s, the last one after an explicit wait
    g_step = [3 * cycles]
    p.reset(0, n)
    own = []
    for c in range(cycles + 1):
        step = 2 * cycles + c + 1
        if c < cycles:
            g.arm(step)
            g.chain(step - 1 if c else 0)
            p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_rec.data_ptr(), stream=st.cuda_stream)
            own.append(d_rec.cpu().numpy().copy())
        else:
            g.wait(step - 1, stream=st.cuda_stream)
        if c:
            host = np.zeros((world * n, 128), np.uint8)
            assert p.lib.dp_memcpy_d2h(p.ctx, abi.ptr(host), C.c_void_p(g.buffer(step - 1)), C.c_size_t(host.nbytes), C.c_void_p(st.cuda_stream)) == 0
            assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
            got.append((host.copy(), own[c - 1]))
    # deferred: the launch of step s forwards / flags / awaits step s-1 as its warps start (dp_gather_arm_deferred); records go
    # to alternating buffers and, in a second round, to ONE buffer (a warp forwards its scene's old record before overwriting it);
    # the last step is pushed by dp_gather_flush
    for bufs, lag in ((2, 1), (1, 1), (2, 2)):
        g.set_lag(lag)                                       # lag 2: the launch of step s awaits step s-2 (a whole step of slack)
        p.reset(0, n)
        d_r = [torch.empty((n, 128), dtype=torch.uint8, device=dev) for _ in range(bufs)]
        own = []
        base = g_step[0]

        def read(step):
            host = np.zeros((world * n, 128), np.uint8)
            assert p.lib.dp_memcpy_d2h(p.ctx, abi.ptr(host), C.c_void_p(g.buffer(step)), C.c_size_t(host.nbytes), C.c_void_p(st.cuda_stream)) == 0
            assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
            got.append((host.copy(), own[step - base - 1]))

        for c in range(cycles):
            g.arm_deferred(base + c + 1)
            p.cycle_dev(n, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), d_r[c % bufs].data_ptr(), stream=st.cuda_stream)
            own.append(d_r[c % bufs].cpu().numpy().copy())
            if c >= lag:
                read(base + c + 1 - lag)                     # complete as soon as this launch is
        g.flush(stream=st.cuda_stream)
        for k in range(lag):
            read(base + cycles - lag + 1 + k)
        g_step[0] = base + cycles
    g.set_lag(1)
    # ... and through the pipelined host-pointer calls (dp_cycle_submit / dp_cycle_wait, two in flight)
    p.reset(0, n)
    recs = [torch.empty((n, 128), dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(n) for _ in range(2)]
    base = g_step[0]
    own, pend = {}, []
    for c in range(cycles):
        if len(pend) == 2:
            p.wait(); k = pend.pop(0); own[k] = recs[k & 1].view(np.uint8).reshape(n, 128).copy()
        g.arm_deferred(base + c + 1)
        p.submit(Hh[c], OXh[c], OYh[c], recs[c & 1])
        pend.append(c)
    while pend:
        p.wait(); k = pend.pop(0); own[k] = recs[k & 1].view(np.uint8).reshape(n, 128).copy()
    g.flush(stream=st.cuda_stream)
    assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
    for c in range(cycles - 4, cycles):                      # depth 4: the last four steps are still in the buffers
        host = np.zeros((world * n, 128), np.uint8)
        assert p.lib.dp_memcpy_d2h(p.ctx, abi.ptr(host), C.c_void_p(g.buffer(base + c + 1)), C.c_size_t(host.nbytes), C.c_void_p(st.cuda_stream)) == 0
        assert p.lib.dp_stream_sync(p.ctx, C.c_void_p(st.cuda_stream)) == 0
        got.append((host.copy(), own[c]))
    q.put((rank, got))
    conn.recv()                                              # keep the mapping alive until the peer has finished too
    g.close(); p.close()


@pytest.mark.parametrize("n_obs", [10, 40])
def test_two_processes_gather_through_cuda_ipc(n_obs):
    world, n, cycles = 2, 96, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe() for _ in range(world)]
    procs = [ctx.Process(target=worker, args=(r, world, pipes[r][1], n, n_obs, cycles, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    handles = [pipes[r][0].recv() for r in range(world)]
    for r in range(world):
        pipes[r][0].send(handles)
    res = dict(q.get(timeout=300) for _ in range(world))
    for r in range(world):
        pipes[r][0].send("done")
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    bad = []
    for c in range(len(res[0])):
        own = [res[r][c][1] for r in range(world)]           # what each rank computed this step
        assert own[0].any()
        for r in range(world):
            gathered = res[r][c][0].reshape(world, n, 128)
            for k in range(world):
                if not np.array_equal(gathered[k], own[k]):
                    bad.append((c, r, k, int((gathered[k] != own[k]).any(axis=1).sum())))
    assert not bad, "(entry, rank, slice of rank, differing records): %s" % bad
