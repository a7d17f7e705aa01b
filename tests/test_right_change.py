"""The right lane change (behaviour 3) of the rule tree: directed scene families (scenes.Directed) that reach it, checked
(1) oracle restatement vs the UNMODIFIED reference, with per-branch hit counters of the oracle, (2) CUDA vs oracle.

Sites that set `behavior = 3` in the reference (Decision.cpp): :1108 navigation-driven, :1382 obstacle-motivated on an
attribute-2 lane, :1618 and :1711 inside the attribute-3 ("both sides") branch.  The last two are DEAD CODE in the
reference itself, for two independent reasons, and the counters prove that their guards are reached but never pass:
  * the right neighbour lane is only loaded when the attribute is exactly 2 (Decision.cpp:636), so with attribute 3 the
    gaps RF / RR are the zeroed defaults and `Path_Obs_RF.dis_lng > Path_Obs_F.dis_lng + 10` is `0 > gF + 10`: false;
  * the signal timer is clamped to 2000 right before the strict test `leftlight_time > 2000` (Decision.cpp:1603-1616,
    1695-1708)."""
import numpy as np
import pytest

from conftest import assert_records_equal, same
from test_oracle_vs_ref import CALL, CARRY, REC

FAMILIES = [("right_obstacle", 0, 512, 40), ("right_nav", 0, 512, 70)]


@pytest.mark.parametrize("family,seed0,n,cycles", FAMILIES)
def test_directed_families_oracle_equals_reference(oracle, reference, the_map, family, seed0, n, cycles):
    from dmpp_b200 import scenes
    ep = scenes.Directed(the_map, np.arange(seed0, seed0 + n), family=family, cycles=cycles)
    H, OX, OY = ep.all_cycles()
    oracle.branch_hits()
    a = oracle.run(H, OX, OY, exhaustive=False, threads=4)
    hits = oracle.branch_hits()
    b = reference.run(H, OX, OY)
    assert b["msgbox"] == 0
    clean = a["trace"]["ub_hits"] == 0          # cycles where the reference reads out of bounds (Planning.cpp:1003-1006) are excluded
    assert clean.mean() > 0.99
    assert_records_equal(a["rec"], b["rec"], REC, mask=clean, what="record")
    assert np.array_equal(a["n_calls"], b["n_calls"])
    assert_records_equal(a["calls"], b["calls"], CALL, what="SearchObstacle call log")
    assert (same(a["path_xy"], b["path_xy"]) | ~clean[..., None, None]).all()
    assert_records_equal(a["carry"], b["carry"], CARRY, what="final state")
    assert same(a["last_path"], b["last_path"]).all()
    # the family reaches what it was written for
    b3 = int((a["rec"]["behavior"] == 3).sum())
    assert b3 >= 1000, b3
    if family == "right_obstacle":
        assert hits["b3_obs_1382"] >= 100 and hits["aim_right_473"] >= 1000 and hits["aim_right_walk"] >= 300, hits
        assert set(np.unique(a["rec"]["behavior_to_dlg"]).tolist()) >= {6, 9}
    else:
        assert hits["b3_nav_1108"] >= 100, hits
        assert set(np.unique(a["rec"]["behavior_to_dlg"]).tolist()) >= {3, 9, 12}
    assert hits["b3_both_1618"] == 0 and hits["b3_both_1711"] == 0


def test_attribute3_right_sites_are_dead_code(oracle, the_map):
    """the guards of Decision.cpp:1596 / :1688 are entered, their `behavior = 3` commits never are (see the module docstring)"""
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(4096), cycles=25)
    oracle.branch_hits()
    oracle.run(*ep.all_cycles(), exhaustive=False, threads=8, paths=False, calls=False, trace=False)
    hits = oracle.branch_hits()
    assert hits["enter_1596"] >= 100, hits
    assert hits["b3_both_1618"] == 0 and hits["b3_both_1711"] == 0, hits


@pytest.mark.gpu
@pytest.mark.parametrize("family,seed0,n,cycles", [("right_obstacle", 10_000, 2048, 40), ("right_nav", 20_000, 2048, 70)])
def test_cuda_right_change(oracle, the_map, family, seed0, n, cycles):
    from dmpp_b200 import scenes
    from dmpp_b200.planner import Planner
    from test_gpu_parity import check
    ep = scenes.Directed(the_map, np.arange(seed0, seed0 + n), family=family, cycles=cycles)
    H, OX, OY = ep.all_cycles()
    want = oracle.run(H, OX, OY, exhaustive=True, threads=8)
    p = Planner(n, OX.shape[2])
    p.upload_map(the_map)
    got = p.run_episodes(H, OX, OY)
    p.close()
    check(got, want, family)
    assert int((got["rec"]["behavior"] == 3).sum()) >= 1000
