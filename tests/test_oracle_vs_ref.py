"""The re-entrant CPU restatement (oracle/planner_oracle.cpp) against the UNMODIFIED reference objects
(oracle/_ref/libref.so: /root/reference/Decision.cpp + Planning.cpp compiled where they lie, driven
through their own thread loops).  Everything the reference publishes must be bit-identical."""
import numpy as np
import pytest

from conftest import assert_records_equal, same

REC = ["velocity_expect", "path_lat_dis", "path_dir_err", "remain_dis", "mindist_lon", "brakespeed", "des_acc", "radius",
       "aim_x", "aim_y", "aim_dir", "aim_id", "behavior", "target_roadnum", "target_lanenum", "light", "behavior_to_dlg",
       "afresh_cause", "path_near_id", "path_front_near_id", "n_traj", "afresh_planning", "acc_flag", "cnt"]
CALL = ["lat_min", "lat_max", "dis_lat", "dis_lng", "n_path", "ob_index", "pathid", "found"]
CARRY = ["leftlight_time", "rightlight_time", "velocity_expect", "aim_x", "aim_y", "aim_dir", "aim_id", "obsavoid_time",
         "no_obsavoid_time", "frontobs_time", "plan_his_behavior", "path_near_id", "behavior", "target_roadnum",
         "target_lanenum", "light_status", "behavior_to_dlg", "his_behavior", "his_target_lanenum", "his_light_status",
         "lanechg_status", "obsavoid_status"]


@pytest.mark.parametrize("kind,seed0,n,cycles,n_obs", [
    ("highway", 0, 768, 25, 10),        # BASELINE config 1/2 shape
    ("highway", 424242, 96, 40, 50),
    ("highway", 9000, 64, 20, 1),
    ("junction", 70000, 192, 80, 10),   # pos 0 -> 1 -> 2 -> 0
    ("junction", 81000, 32, 60, 40),
])
def test_restatement_equals_unmodified_reference(oracle, reference, the_map, kind, seed0, n, cycles, n_obs):
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(seed0, seed0 + n), cycles=cycles, kind=kind, n_obs=n_obs)
    H, OX, OY = ep.all_cycles()
    a = oracle.run(H, OX, OY, exhaustive=False, threads=4)
    b = reference.run(H, OX, OY)
    assert b["msgbox"] == 0
    # cycles where the reference itself reads out of bounds (Planning.cpp:1003-1006) are defined by the oracle
    # and excluded here; they must stay rare
    clean = a["trace"]["ub_hits"] == 0
    assert clean.mean() > 0.999
    assert_records_equal(a["rec"], b["rec"], REC, mask=clean, what="record")
    assert np.array_equal(a["n_calls"], b["n_calls"])
    assert_records_equal(a["calls"], b["calls"], CALL, what="SearchObstacle call log")
    assert (same(a["path_xy"], b["path_xy"]) | ~clean[..., None, None]).all()
    assert (same(a["path_ll"], b["path_ll"]) | ~clean[..., None, None]).all()
    assert_records_equal(a["carry"], b["carry"], CARRY, what="final state")
    assert same(a["last_path"], b["last_path"]).all()
    assert a["traj"] == b["traj"]


def test_scene_mix_is_not_trivial(oracle, the_map):
    """the scripted episodes reach every branch family of the rule tree the map allows"""
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(2048), cycles=25)
    a = oracle.run(*ep.all_cycles(), exhaustive=False, threads=8)
    assert set(np.unique(a["rec"]["behavior"]).tolist()) >= {1, 2, 4, 5}
    assert set(np.unique(a["rec"]["behavior_to_dlg"]).tolist()) >= {1, 2, 3, 4, 5, 6, 8, 9, 11, 12}
    assert set(np.unique(a["rec"]["afresh_cause"]).tolist()) == {0, 1, 2, 4}
    assert a["ub_hits"] == 0


def test_oracle_is_thread_count_invariant(oracle, the_map):
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(5000, 5200), cycles=12)
    H, OX, OY = ep.all_cycles()
    a = oracle.run(H, OX, OY, threads=1)
    b = oracle.run(H, OX, OY, threads=7)
    assert a["rec"].tobytes() == b["rec"].tobytes() and a["trace"].tobytes() == b["trace"].tobytes()
