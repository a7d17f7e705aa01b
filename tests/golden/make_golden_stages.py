"""Golden fixtures of the stages either side of the cycle, from the UNMODIFIED reference (oracle/_ref/libref.so; exists only where
/root/reference is mounted).  Run from the repo root:  python tests/golden/make_golden_stages.py

  stages/closed_loop_24.npz  the reference's own thread loops run CLOSED-LOOP (oracle/ref_closed_loop.cpp: the frozen world step of
                             oracle/world_spec.cpp between its cycles): records, logged headers / obstacle points, final world
  stages/frames_*.npz        the PlanningOut / PlanningStatus objects its Planning thread publishes every cycle, in the frame layout
  stages/v2x_2048.npz        the flags of its V2XEventDecision / V2XConstructionEventTemporal on seeded events
Inputs are not stored: they are pure functions of the seeds (scenes.World / Episodes / v2x_events); a checksum guards against drift."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from oracle import binding  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stages")
CLOSED = (4000, 24, 50, 10)                       # first seed, worlds, cycles, obstacles
FRAMES = {"frames_highway_16": ("highway", 200, 16, 25), "frames_junction_8": ("junction", 71000, 8, 60)}
V2X = (0, 2048, 11)                               # first seed, scenes, event seed


def checksum(*arrays):
    return np.array([np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).astype(np.uint64).sum() for a in arrays], np.uint64)


def closed_loop_inputs(m):
    s0, n, cycles, n_obs = CLOSED
    w = scenes.World(m, np.arange(s0, s0 + n), n_obs=n_obs)
    return w, cycles


def frames_inputs(m, name):
    kind, s0, n, cycles = FRAMES[name]
    return scenes.Episodes(m, np.arange(s0, s0 + n), cycles=cycles, kind=kind).all_cycles()


def v2x_inputs(m):
    s0, n, seed = V2X
    h = scenes.Episodes(m, np.arange(s0, s0 + n), cycles=1).hdr(0)
    return (h,) + scenes.v2x_events(m, h, seed=seed)


def main():
    m = scenes.Map()
    ref = binding.Reference(); ref.set_map(m)
    orc = binding.Oracle(); orc.set_map(m)          # (parameters and the UB mask only)
    os.makedirs(HERE, exist_ok=True)
    w, cycles = closed_loop_inputs(m)
    o = ref.run_closed_loop(orc.params, orc.world_params(), w.hdr, w.agents, cycles, paths=True)
    ub = orc.run_closed_loop(w.hdr, w.agents, cycles)["ub_scene"]
    np.savez_compressed(os.path.join(HERE, "closed_loop_24.npz"), rec=o["rec"], hdr_log=o["hdr_log"], obs_log_x=o["obs_log_x"],
                        obs_log_y=o["obs_log_y"], hdr=o["hdr"], agents=o["agents"], last_path=o["last_path"], ub_scene=ub,
                        inputs=checksum(w.hdr, w.agents))
    print("closed_loop_24", o["rec"].shape, "scenes with reference UB:", int((ub > 0).sum()))
    for name in FRAMES:
        H, OX, OY = frames_inputs(m, name)
        o = ref.run_with_frames(H, OX, OY)
        clean = orc.run(H, OX, OY, exhaustive=False)["trace"]["ub_hits"] == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ctrl=o["ctrl"], status=o["status"], clean=clean, inputs=checksum(H, OX, OY))
        print(name, o["ctrl"].shape)
    h, v, wl, wg = v2x_inputs(m)
    f0, f1 = ref.v2x_event(h, v, wl, wg, 0), ref.v2x_event(h, v, wl, wg, 1)
    u0, u1 = orc.v2x_event(h, v, wl, wg, 0)["ub"], orc.v2x_event(h, v, wl, wg, 1)["ub"]
    keep = ["light_flag", "construction_flag", "pedestrian_flag"]
    np.savez_compressed(os.path.join(HERE, "v2x_2048.npz"), mode0=np.stack([f0[k] for k in keep]), mode1=np.stack([f1[k] for k in keep]),
                        ub0=u0, ub1=u1, inputs=checksum(h, v, wl, wg))
    print("v2x_2048", int(f0["pedestrian_flag"].sum()), int(f0["construction_flag"].sum()), int(f1["construction_flag"].sum()))


if __name__ == "__main__":
    main()
