"""Regenerates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libref.so, which only
exists where /root/reference is mounted).  Run from the repo root:  python tests/golden/make_golden.py

Fixtures are small .npz files: the inputs are NOT stored (they are a pure function of the seeds,
decision-making-and-path-planning_b200/scenes.py); the outputs of the reference are."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import scenes  # noqa: E402
from oracle import binding  # noqa: E402

CASES = {
    # name: (kind, first seed, scenes, cycles, obstacles)
    "highway_10": ("highway", 0, 48, 25, 10),
    "highway_40": ("highway", 31337, 16, 30, 40),
    "junction_10": ("junction", 555000, 24, 80, 10),
    # directed families that reach the right lane change (behaviour 3), scenes.Directed
    "right_obstacle_10": ("right_obstacle", 0, 32, 40, 10),
    "right_nav_10": ("right_nav", 0, 32, 70, 10),
}


def episodes(m, kind, seeds, cycles, n_obs):
    if kind.startswith("right_"):
        return scenes.Directed(m, seeds, family=kind, cycles=cycles, n_obs=n_obs)
    return scenes.Episodes(m, seeds, cycles=cycles, kind=kind, n_obs=n_obs)


def sig(a, axis_from):
    """order-sensitive, bit-exact 64-bit signature of the trailing axes of a float64/structured array"""
    a = np.ascontiguousarray(a)
    lead = a.shape[:axis_from]
    b = np.frombuffer(a.tobytes(), np.uint64).reshape(int(np.prod(lead)) if lead else 1, -1)
    w = (2 * np.arange(b.shape[1], dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        return (b * w).sum(axis=1, dtype=np.uint64).reshape(lead)


def main():
    m = scenes.Map()
    ref = binding.Reference()
    ref.set_map(m)
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (kind, s0, n, cyc, nobs) in CASES.items():
        ep = episodes(m, kind, np.arange(s0, s0 + n), cyc, nobs)
        H, OX, OY = ep.all_cycles()
        o = ref.run(H, OX, OY)
        # checksum of the inputs so that a drift of the generator is detected, not silently absorbed
        inp = np.array([np.frombuffer(H.tobytes(), np.uint8).astype(np.uint64).sum(), OX.sum(), OY.sum()])
        np.savez_compressed(os.path.join(here, name + ".npz"), rec=o["rec"], path_sig=sig(o["path_xy"], 2),
                            path_ll_sig=sig(o["path_ll"], 2), path_xy_first=o["path_xy"][:, :2],
                            n_calls=o["n_calls"], calls_sig=sig(o["calls"], 2), calls_first=o["calls"][:, :2], carry=o["carry"],
                            last_path_sig=sig(o["last_path"], 1), inputs=inp, meta=np.array([s0, n, cyc, nobs]), kind=np.array(kind))
        print(name, "written", o["rec"].shape)


if __name__ == "__main__":
    main()
