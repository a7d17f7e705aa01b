#!/usr/bin/env python
"""bench.py -- headline benchmark of the Decision/Planning hot path (BASELINE.json metric:
trajectories scored / s and plan cycles / s).

Workload (BASELINE config 2): 4096 independent synthetic highway scenes per GPU, ego + 10
vehicles, reference-derived default candidate set, 25-cycle scripted episodes.  One "step" = one
Decision+Planning cycle for every scene of the batch (the Decision and the Planning launch of dp_cycle_kernel, overlapped).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one rank per GPU)
  python bench.py --impl reference [...]                       the reference's own CPU code (oracle/_ref)

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each number is obtained.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import abi, scenes  # noqa: E402

SCENES = 4096
N_OBS = 10
EPISODE = 25
METRIC = "trajectories_scored_per_s"
UNIT = "trajectories/s"


def workload_desc(world):
    return {"workload": ("config2" if SCENES == 4096 else "config4 shard") + ": %d highway scenes/GPU x default candidate set (6 lane regions + 2K avoid offsets + 1 local path), "
                        "ego + %d vehicles, %d-cycle scripted episodes" % (SCENES, N_OBS, EPISODE),
            "scenes_per_gpu": SCENES, "obstacles": N_OBS, "episode_cycles": EPISODE,
            "parallelism": "scenes sharded %d-way, no data-path collective in the compute phase; when N>1 every rank's plan records are "
                           "gathered on every rank each step: GATHER_KIND; the gather of step i is finished beside the kernels of "
                           "step i+1 (triple-buffered), the last one is exposed and counted" % world,
            "l2": "256 MiB buffer written between timed steps (L2 flush); inputs resident in HBM for `value`"}


def alg_flops(pts, ntraj, n_obs):
    """F(P,N) = 18 P + N (5 P + 12) summed over the trajectories actually scored."""
    return 18.0 * pts + n_obs * (5.0 * pts + 12.0 * ntraj)


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref/libref.so), one process per core
# ------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(seed0, n, cycles):
    from oracle import binding
    m = scenes.Map()
    if binding.Reference.available():
        r = binding.Reference(); kind = "reference"
    else:
        r = binding.Oracle(); kind = "port"
    r.set_map(m)
    ep = scenes.Episodes(m, np.arange(seed0, seed0 + n), cycles=cycles, n_obs=N_OBS)
    _W.update(r=r, kind=kind, data=ep.all_cycles())
    return kind


def _worker_run(_):
    H, OX, OY = _W["data"]
    if _W["kind"] == "reference":
        o = _W["r"].run(H, OX, OY, paths=False, calls=False)
    else:
        o = _W["r"].run(H, OX, OY, paths=False, calls=False, trace=False, exhaustive=False, threads=1)
    return int(o["traj"]), float(o["seconds"]), int(H.size)


class CpuArm:
    """P worker processes, each owning a private copy of the (non re-entrant) reference and a scene slice."""

    def __init__(self, total_scenes, cycles, procs):
        import multiprocessing as mp
        self.procs = procs
        self.ctx = mp.get_context("spawn")
        per = (total_scenes + procs - 1) // procs
        self.pools = [self.ctx.Pool(1) for _ in range(procs)]
        kinds = [p.apply_async(_worker_init, (10_000_000 + i * per, per, cycles)) for i, p in enumerate(self.pools)]
        self.kind = kinds[0].get()
        for k in kinds:
            k.get()

    def step(self):
        t0 = time.perf_counter()
        res = [p.apply_async(_worker_run, (0,)) for p in self.pools]
        res = [r.get() for r in res]
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res), dt, sum(r[2] for r in res)

    def close(self):
        for p in self.pools:
            p.terminate()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    arm = CpuArm(min(SCENES, 4096), EPISODE, cores)
    for _ in range(max(1, min(args.warmup, 2))):
        arm.step()
    traj = 0; secs = 0.0; cyc = 0
    steps = max(1, min(args.steps, 20))
    for _ in range(steps):
        t, dt, c = arm.step()
        traj += t; secs += dt; cyc += c
    arm.close()
    val = traj / secs
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": secs / cyc * min(SCENES, 4096) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_desc(1),
            "plan_cycles_per_s": cyc / secs,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": arm.kind, "per_core": val / max(cores, 1),
                             "sample": "%d steps, each = %d scenes x %d cycles (whole episodes), one process per core running "
                                       "the unmodified Decision.cpp/Planning.cpp objects" % (steps, min(SCENES, 4096), EPISODE)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "ms_per_step is normalised to one plan cycle of %d scenes" % min(SCENES, 4096)}
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop = False; self.sm = []; self.reasons = set(); self.max_sm = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            while not self.stop:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dmpp_b200.planner import Planner

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the record gather runs beside the next step's kernels: keep its footprint to a few CTAs so that the one-wave
        # Decision launch (1024 CTAs on 1036 slots) still fits
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, args.warmup
    m = scenes.Map()
    ep = scenes.Episodes(m, np.arange(rank * SCENES, (rank + 1) * SCENES), cycles=EPISODE, n_obs=N_OBS)
    H, OX, OY = ep.all_cycles()
    planner = Planner(max_scenes=SCENES, max_obs=N_OBS, device=local_rank)
    planner.upload_map(m)

    # ---- untimed replay with the trace on: trajectories and path points scored per cycle (deterministic) ----
    rep = planner.run_episodes(H, OX, OY, trace=True, paths=False)
    traj_c = rep["rec"]["n_traj"].astype(np.int64).sum(axis=1)
    pts_c = rep["trace"]["pts_scored"].astype(np.int64).sum(axis=1)
    flops_c = alg_flops(pts_c.astype(np.float64), traj_c.astype(np.float64), N_OBS)

    # ---- inputs resident in HBM ----
    d_hdr = torch.from_numpy(H.view(np.uint8).reshape(EPISODE, SCENES, 128)).to(dev)
    d_ox = torch.from_numpy(OX).to(dev)
    d_oy = torch.from_numpy(OY).to(dev)
    d_recs = [torch.empty((SCENES, 128), dtype=torch.uint8, device=dev) for _ in range(3)]   # plan records, rotated per step
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(W + K)]
    # N > 1: the records of every rank are gathered on every rank each step.  The gather is fused into the Planning launch:
    # each rank's kernel stores its finished records straight into its slice of every peer's gathered buffer over NVLink
    # (torch symmetric memory gives the peer mappings, dp_set_record_mirrors hands them to the kernel); what is left of the
    # collective is a barrier, run on a side stream beside the next step's kernels.  Three gathered buffers rotate so that a
    # step never overwrites records a peer may still be reading.
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    gath = hdl = None
    gather_kind = "none"
    if world > 1:
        try:
            if os.environ.get("DP_BENCH_NO_SYMM"):
                raise RuntimeError("disabled by DP_BENCH_NO_SYMM")
            import torch.distributed._symmetric_memory as symm
            gath = [symm.empty((world * SCENES, 128), dtype=torch.uint8, device=dev) for _ in range(3)]
            hdl = [symm.rendezvous(g, dist.group.WORLD) for g in gath]
            for g in gath:
                g.zero_()
            gather_kind = "peer stores from the Planning launch (NVLink, symmetric memory) + barrier on a side stream"
        except Exception as e:  # noqa: BLE001 -- no peer mapping on this box: fall back to NCCL's all_gather
            print("symmetric memory unavailable (%s): using all_gather_into_tensor" % e, file=sys.stderr)
            gath = [torch.empty((world * SCENES, 128), dtype=torch.uint8, device=dev) for _ in range(3)]
            hdl = None
            gather_kind = "all_gather_into_tensor on a side stream"

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tail_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def finish_gather(i, rec_i):
        """what is left of step i's gather after its kernels: a barrier (peer stores) or the whole all_gather (fallback)"""
        if hdl is not None:
            hdl[i % 3].barrier()
        else:
            dist.all_gather_into_tensor(gath[i % 3], rec_i)

    def dev_loop(first, count):
        """N > 1: the gather of step i is finished on a side stream concurrently with the kernels of step i+1 (it is
        released by the start event of step i+1, so it cannot hide inside the untimed L2 flush); a step is complete when
        its kernels AND the previous step's gather are done; the last one is timed on its own."""
        pending = None                                       # (step, records) whose gather is not finished yet
        for i in range(first, first + count):
            c = i % EPISODE
            if c == 0:
                torch.cuda.synchronize()
                planner.reset(0, SCENES)
            flush.zero_()
            rec_i = d_recs[i % 3]
            if hdl is not None:                              # this step's records go to slice `rank` of every rank's buffer i % 3
                planner.set_record_mirrors([p + rank * SCENES * 128 for p in hdl[i % 3].buffer_ptrs])
            ev[i][0].record(stream)
            if pending is not None:
                comm.wait_event(ev[i][0])
                with torch.cuda.stream(comm):
                    finish_gather(*pending)
                    gdone = torch.cuda.Event()
                    gdone.record(comm)
            planner.cycle_dev(SCENES, d_hdr[c].data_ptr(), d_ox[c].data_ptr(), d_oy[c].data_ptr(), rec_i.data_ptr(),
                              stream=stream.cuda_stream)
            ev[i][1].record(stream)
            if pending is not None:
                stream.wait_event(gdone)
            ev[i][2].record(stream)
            pending = (i, rec_i) if world > 1 else None
        if world > 1:
            tail_ev[0].record(stream)
            finish_gather(*pending)
            tail_ev[1].record(stream)
        return pending

    barrier()
    dev_loop(0, W)
    barrier()
    l0 = planner.launch_count()
    last = dev_loop(W, K)
    barrier()
    if world > 1:                                            # the gathered buffer of the last step holds every rank's records
        g = gath[last[0] % 3].view(world, SCENES, 128)
        assert torch.equal(g[rank], last[1]), "own slice of the gathered records differs"
        chk = torch.tensor([float(last[1][:, :].sum(dtype=torch.int64).item())], dtype=torch.float64, device=dev)
        allchk = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        for r in range(world):
            assert float(g[r].sum(dtype=torch.int64).item()) == float(allchk[r].item()), "gathered slice of rank %d differs" % r
        planner.set_record_mirrors([])
    launches = planner.launch_count() - l0 - sum(1 for i in range(W, W + K) if i % EPISODE == 0)
    kern_ms = np.array([ev[i][0].elapsed_time(ev[i][1]) for i in range(W, W + K)])
    step_ms = np.array([ev[i][0].elapsed_time(ev[i][2]) for i in range(W, W + K)])
    tail_ms = tail_ev[0].elapsed_time(tail_ev[1]) if world > 1 else 0.0
    cyc_idx = np.array([i % EPISODE for i in range(W, W + K)])
    total_ms = torch.tensor([step_ms.sum() + tail_ms], dtype=torch.float64, device=dev)
    traj = torch.tensor([float(traj_c[cyc_idx].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(traj, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) * 1e-3
    value = float(traj.item()) / total_s

    # ---- end to end through the host-pointer C ABI: pinned host inputs, H2D + kernel + D2H every step ----
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    Hh = pin(H.view(np.uint8).reshape(EPISODE, SCENES, 128)).view(abi.scene_hdr).reshape(EPISODE, SCENES)
    OXh, OYh = pin(OX), pin(OY)
    rec_h = torch.empty((SCENES, 128), dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(SCENES)
    out = {"rec": rec_h}
    e2e_t = np.zeros(W + K)
    barrier()
    for i in range(W + K):
        c = i % EPISODE
        if c == 0:
            planner.reset(0, SCENES)
        if i == W:
            barrier()
        t0 = time.perf_counter()
        planner.cycle(Hh[c], OXh[c], OYh[c], out=out)
        e2e_t[i] = time.perf_counter() - t0
    barrier()
    e2e_s = torch.tensor([e2e_t[W:].sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_sync_val = float(traj.item()) / float(e2e_s.item())
    assert int(rec_h["n_traj"].astype(np.int64).sum()) == int(traj_c[(W + K - 1) % EPISODE]), "e2e result differs from replay"

    # ---- the same, pipelined: dp_cycle_submit / dp_cycle_wait, two cycles in flight.  Every step still moves its own
    #      inputs host->device and its records device->host inside the timed region; the region ends after the last wait. ----
    recs2 = [rec_h, torch.empty((SCENES, 128), dtype=torch.uint8).pin_memory().numpy().view(abi.plan_record).reshape(SCENES)]

    CH = 32768                                               # dp_cycle_submit takes at most 32768 scenes per call

    def pipe_loop(first, count):
        got = 0; pend = []                                   # record slices in flight (at most two)
        for i in range(first, first + count):
            c = i % EPISODE
            if c == 0:
                while pend:
                    planner.wait(); got += int(pend.pop(0)["n_traj"].sum(dtype=np.int64))
                planner.reset(0, SCENES)
            for s0 in range(0, SCENES, CH):
                s1 = min(SCENES, s0 + CH)
                if len(pend) == 2:
                    planner.wait(); got += int(pend.pop(0)["n_traj"].sum(dtype=np.int64))
                planner.submit(Hh[c, s0:s1], OXh[c, s0:s1], OYh[c, s0:s1], recs2[i & 1][s0:s1], first=s0)
                pend.append(recs2[i & 1][s0:s1])
        while pend:
            planner.wait(); got += int(pend.pop(0)["n_traj"].sum(dtype=np.int64))
        return got

    barrier()
    pipe_loop(0, W)
    barrier()
    t0 = time.perf_counter()
    got = pipe_loop(W, K)
    pipe_t = time.perf_counter() - t0
    barrier()
    assert got == int(traj_c[cyc_idx].sum()), "pipelined e2e results differ from replay"
    e2e_p = torch.tensor([pipe_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_p, op=dist.ReduceOp.MAX)
    e2e_val = float(traj.item()) / float(e2e_p.item())
    sampler.stop = True
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp64_tf, fp32_tf = planner.measure_fma_peak()
    achieved = float(flops_c[cyc_idx].sum()) / (kern_ms.sum() * 1e-3) / 1e12
    bytes_scene = 128 + N_OBS * 16 + 2 * 128 + 3200 + 128        # hdr + obstacles + carry r/w + last path read + record
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:   # DRAM bytes per launch of dp_cycle_kernel from the committed `ncu --set full` capture of tools/profile_cycle.py
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_r1.json")))["dp_cycle_kernel"]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": float(step_ms.mean()), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {k: (v.replace("GATHER_KIND", gather_kind) if isinstance(v, str) else v)
                                                        for k, v in workload_desc(world).items()},
        "plan_cycles_per_s": world * SCENES * K / total_s,
        "trajectories_per_step_per_gpu": float(traj_c[cyc_idx].mean()),
        "cycle_latency_ms": {"p50": float(np.percentile(step_ms, 50)), "p99": float(np.percentile(step_ms, 99)),
                             "max": float(step_ms.max()), "what": "device time of one Decision+Planning cycle of %d scenes" % SCENES},
        "e2e": {"value": e2e_val, "unit": UNIT,
                "h2d_bytes_per_step": SCENES * (128 + 2 * N_OBS * 8), "d2h_bytes_per_step": SCENES * 128,
                "ms_per_step": float(e2e_p.item()) / K * 1e3, "plan_cycles_per_s": world * SCENES * K / float(e2e_p.item()),
                "api": "dp_cycle_submit + dp_cycle_wait (host pointers, pinned; two cycles in flight: the inputs of step i+1 cross "
                       "PCIe while step i computes; every step's records are read on the host; timed region ends after the last wait)",
                "synchronous": {"value": e2e_sync_val, "ms_per_step": float(e2e_t[W:].mean() * 1e3),
                                "plan_cycles_per_s": world * SCENES * K / float(e2e_s.item()),
                                "api": "dp_cycle_batch (host pointers, pinned; returns when the records are valid)"}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "kernel": "dp_cycle_kernel", "achieved": achieved, "peak": fp64_tf, "unit": "TFLOP/s",
                     "frac": achieved / fp64_tf if fp64_tf else None, "traffic": traffic,
                     "peak_source": "FP64 FMA micro-benchmark run in this process (dp_measure_fma_peak); MEASURED_PEAKS.json has no "
                                    "CUDA-core FP64 figure. fp32 FMA peak measured the same way: %.1f TFLOP/s" % fp32_tf,
                     "algorithmic_flops_per_launch": float(flops_c[cyc_idx].mean()),
                     "kernel_ms": float(kern_ms.mean()),
                     "hbm": {"achieved": SCENES * bytes_scene / (kern_ms.mean() * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "algorithmic_bytes_per_launch": SCENES * bytes_scene,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
        "clocks": sampler.result(),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = host_cores()
            arm = CpuArm(min(SCENES, 4096), EPISODE, cores)
            arm.step()
            t_traj = 0; t_s = 0.0; n = 0
            while t_s < 4.0 and n < 40:
                t, dt, _ = arm.step()
                t_traj += t; t_s += dt; n += 1
            arm.close()
            line["cpu_baseline"] = {"value": t_traj / t_s, "unit": UNIT, "cores": cores, "kind": arm.kind, "per_core": t_traj / t_s / max(cores, 1),
                                    "sample": "%d passes over %d scenes x %d cycles, one process per core (unmodified reference "
                                              "objects when kind=reference)" % (n, min(SCENES, 4096), EPISODE)}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    emit(line)
    planner.close()
    if world > 1:
        dist.destroy_process_group()


_OUT_FD = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner ...) was
    redirected to stderr in main()"""
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def main():
    global _OUT_FD
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)                                            # C-level and Python-level stdout -> stderr from here on
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scenes", type=int, default=SCENES,
                    help="scenes per GPU (default = BASELINE config 2; 131072 x 8 GPUs = config 4, the 1M-scene sweep)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    globals()["SCENES"] = args.scenes
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
