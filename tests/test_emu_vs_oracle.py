"""The LOGIC of the group kernel (csrc/dp_group.cuh: phase structure, exactly pruned scans, prefix-table walks, rule tree) on
a box without a GPU: the kernel source compiled by g++ with every phase as a loop over the CTA's thread ids (tools/emu,
test infrastructure) against the oracle, bit for bit, on the same case families as the GPU parity suite.  What this cannot
see -- real concurrency, memory ordering, the TMA copy -- is what tests/test_gpu_parity.py is for."""
import numpy as np
import pytest

from test_gpu_parity import check


@pytest.fixture(scope="module")
def emu(the_map):
    from tools.emu.binding import Emu
    e = Emu()
    e.set_map(the_map)
    return e


CASES = [("highway", 0, 768, 25, 10, 14), ("junction", 100000, 128, 80, 10, 16), ("highway", 5000, 96, 25, 1, 1),
         ("highway", 5100, 96, 25, 33, 13), ("junction", 7000, 16, 80, 200, 16), ("right_obstacle", 0, 128, 40, 10, 7),
         ("right_nav", 0, 128, 70, 10, 16)]


@pytest.mark.parametrize("kind,seed0,n,cycles,n_obs,group", CASES)
def test_group_kernel_logic_matches_oracle(emu, oracle, the_map, kind, seed0, n, cycles, n_obs, group):
    from dmpp_b200 import scenes
    seeds = np.arange(seed0, seed0 + n)
    ep = (scenes.Directed(the_map, seeds, family=kind, cycles=cycles, n_obs=n_obs) if kind.startswith("right_")
          else scenes.Episodes(the_map, seeds, cycles=cycles, kind=kind, n_obs=n_obs))
    H, OX, OY = ep.all_cycles()
    want = oracle.run(H, OX, OY, exhaustive=True, threads=4)
    got = emu.run(H, OX, OY, group=group)
    check(got, want, kind)
    fast = emu.run(H, OX, OY, group=group, trace=False, paths=False)   # the untraced path stops its sums early: same records
    assert fast["rec"].tobytes() == got["rec"].tobytes()


@pytest.mark.parametrize("n_obs,n,group", [(50, 64, 14), (200, 16, 16)])
def test_group_kernel_logic_with_predicted_tracks(emu, oracle, the_map, n_obs, n, group):
    """BASELINE config 5: junction search against constant-turn-rate track tiles (T = 400), ~400-point reference paths"""
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(900, 900 + n), cycles=40, kind="urban", n_obs=n_obs)
    H, OX, OY, VX, VY, DTH = ep.all_cycles_tracks()
    want = oracle.run_tracks(H, OX, OY, VX, VY, DTH, ep.TRACK_T, threads=4)
    static = oracle.run(H, OX, OY, threads=4)
    assert (want["trace"]["junction"]["pathid"] != static["trace"]["junction"]["pathid"]).mean() > 0.3   # the tracks matter
    assert np.percentile(want["trace"]["refpath_len"], 50) > 300
    got = emu.run_tracks(H, OX, OY, VX, VY, DTH, ep.TRACK_T, group=group)
    check(got, want, "urban tracks")
