"""Closed-loop episodes (include/dmpp_b200.h section 9) and the output frames (section 8).

CPU side: the restated oracle with the frozen world step (oracle/world_spec.cpp) between its cycles against the UNMODIFIED
reference driven the same way (oracle/ref_closed_loop.cpp), and the frame restatement against the PlanningOut / PlanningStatus
objects the reference's Planning thread itself publishes.  GPU side: dp_run_closed_loop_dev (one CUDA graph per episode) and
dp_pack_frames against the oracle, bit for bit; the inputs the closed loop logs, replayed open loop, give the same records."""
import os

import numpy as np
import pytest

from conftest import assert_records_equal, same
from test_oracle_vs_ref import CARRY, REC

HDR = ["x", "y", "dir", "velocity", "period_ms", "id", "road_num", "lane_num", "pos", "path_num", "out_lane_no", "n_obs"]
GPU_REC = REC + ["mindist_lat", "sweep_index", "ob_index", "ob_pathid", "ob_flag"]
DIR_ERR_TOL = {"path_dir_err": 1e-9}


def world(the_map, seed0, n, n_obs=10, roads=None):
    from dmpp_b200 import scenes
    return scenes.World(the_map, np.arange(seed0, seed0 + n), n_obs=n_obs, roads=roads)


# ---------------------------------------------------------------- CPU: oracle vs the unmodified reference
@pytest.mark.parametrize("seed0,n,cycles,n_obs", [(0, 256, 60, 10), (31000, 48, 120, 24), (777, 64, 40, 1)])
def test_closed_loop_oracle_equals_unmodified_reference(oracle, reference, the_map, seed0, n, cycles, n_obs):
    w = world(the_map, seed0, n, n_obs)
    a = oracle.run_closed_loop(w.hdr, w.agents, cycles, paths=True, threads=4)
    b = reference.run_closed_loop(oracle.params, oracle.world_params(), w.hdr, w.agents, cycles, paths=True)
    assert b["status"] == 0                                 # no message box in the reference
    # scenes in which the reference itself reads out of bounds (Planning.cpp:1003-1006: the oracle defines the value by
    # clamping) are excluded as everywhere; they must stay rare
    clean = a["ub_scene"] == 0
    assert clean.mean() > 0.98
    assert_records_equal(a["rec"][:, clean], b["rec"][:, clean], REC, what="record")
    assert_records_equal(a["hdr_log"][:, clean], b["hdr_log"][:, clean], HDR, what="logged header")
    for k in ("obs_log_x", "obs_log_y", "path_xy"):
        assert same(a[k][:, clean], b[k][:, clean]).all(), k
    assert_records_equal(a["hdr"][clean], b["hdr"][clean], HDR, what="final header")
    assert a["agents"][clean].tobytes() == b["agents"][clean].tobytes()
    assert_records_equal(a["carry"][clean], b["carry"][clean], CARRY, what="final state")
    assert same(a["last_path"][clean], b["last_path"][clean]).all()


def test_closed_loop_world_is_alive(oracle, the_map):
    """the loop is closed: the ego follows its own plans (speed commands change its speed, its map index advances, lane changes
    move it to another lane), the agents move, and the hysteresis of the rule tree shows up over the episode"""
    w = world(the_map, 0, 512)
    a = oracle.run_closed_loop(w.hdr, w.agents, 80, threads=8)
    H = a["hdr_log"]
    lane = H["lane_num"].astype(int)
    ego_id = np.take_along_axis(H["id"], (lane - 1)[..., None], axis=2)[..., 0]
    moving = H["velocity"][0] > 0
    assert (ego_id[-1] >= ego_id[0]).all() and (ego_id[-1] > ego_id[0])[moving].mean() > 0.9
    assert (np.abs(H["velocity"][-1] - H["velocity"][0]) > 1.0).mean() > 0.5          # speed follows the planner's command
    assert (lane[-1] != lane[0]).sum() >= 5                                          # completed lane changes
    assert np.abs(a["obs_log_x"][-1] - a["obs_log_x"][0]).max() > 10.0
    assert set(np.unique(a["rec"]["behavior"]).tolist()) >= {1, 2, 4, 5}
    assert set(np.unique(a["rec"]["afresh_cause"]).tolist()) >= {0, 1, 4}
    assert np.abs(a["rec"]["path_lat_dis"][1:]).max() < 0.5                          # perfect tracking: the ego stays on its path
    t = oracle.run_closed_loop(w.hdr, w.agents, 80, threads=3)
    assert t["rec"].tobytes() == a["rec"].tobytes() and t["hdr"].tobytes() == a["hdr"].tobytes()


def test_world_step_known_answers(oracle, the_map):
    """straight road 1 (heading 0, 0.5 m spacing): positions, indices and speeds that can be written down by hand"""
    from dmpp_b200 import abi
    m = the_map
    h = np.zeros(1, abi.scene_hdr)
    gl = m.lane_index(1, 2)
    off = int(m.lane_pt_off[gl])
    h["x"], h["y"], h["dir"], h["velocity"], h["period_ms"] = m.x[off + 100], m.y[off + 100], 0.0, 36.0, 100.0
    h["road_num"], h["lane_num"], h["n_obs"] = 1, 2, 2
    h["id"][0, :3] = 100
    ag = np.zeros((1, 2), abi.agent)
    ag["lane"] = [[gl, m.lane_index(1, 1)]]
    ag["i"], ag["u"], ag["v"], ag["lat"] = [[140, 90]], [[0.25, 0.0]], [[10.0, 0.0]], [[0.5, -1.0]]
    ox, oy = np.zeros((1, 2)), np.zeros((1, 2))
    oracle.world_step(h, ag, ox, oy)                                                 # place + localise
    assert ox[0, 0] == m.x[off + 140] + 0.25 and oy[0, 0] == m.y[off + 140] + 0.5    # 0.5 m to the LEFT (north on a road heading east)
    assert oy[0, 1] == m.y[int(m.lane_pt_off[ag["lane"][0, 1]]) + 90] - 1.0
    assert h["id"][0, :3].tolist() == [100, 100, 100] and h["lane_num"][0] == 2
    rec = np.zeros(1, abi.plan_record)
    rec["brakespeed"], rec["path_near_id"] = 10.0, 0                                 # 10 m/s = 36 km/h: keep the speed
    lp = np.zeros((1, 2, 200))
    lp[0, 0] = h["x"][0] + 0.25 * np.arange(200)                                     # a straight path east, 0.25 m spacing
    lp[0, 1] = h["y"][0]
    oracle.world_step(h, ag, ox, oy, rec=rec, last_path=lp)
    assert h["velocity"][0] == 36.0 and h["x"][0] == m.x[off + 100] + 1.0 and h["dir"][0] == 0.0   # 10 m/s x 0.1 s
    assert h["id"][0, 1] == 102 and h["lane_num"][0] == 2
    assert ag["i"][0, 0] == 142 and ag["u"][0, 0] == 0.25 and ox[0, 0] == m.x[off + 142] + 0.25    # 10 m/s x 0.1 s = 2 points
    rec["brakespeed"] = 0.0                                                          # brake: |dv| <= 3 m/s^2 x 0.1 s = 1.08 km/h
    oracle.world_step(h, ag, ox, oy, rec=rec, last_path=lp)
    assert h["velocity"][0] == 36.0 - 3.0 * 3.6 * 0.1


def test_world_step_invariants(oracle, the_map):
    """properties of the world step that do not depend on its exact arithmetic: bounded acceleration, forward motion, agents
    that keep their lateral offset from their lane, indices inside the localisation window, a localised ego on the right lane"""
    from dmpp_b200 import abi
    w = world(the_map, 2024, 384)
    cycles = 40
    a = oracle.run_closed_loop(w.hdr, w.agents, cycles, threads=4)
    H = a["hdr_log"]
    wp = oracle.world_params()
    dt = H["period_ms"][0] / 1000.0
    ids = H["id"].astype(int)
    lane = H["lane_num"].astype(int)
    ego_id = np.take_along_axis(ids, (lane - 1)[..., None], axis=2)[..., 0]
    at_end = ego_id >= 2000 - wp.end_margin                                         # (an ego at the end of its lane is stopped outright)
    dv = np.abs(np.diff(H["velocity"], axis=0))
    assert (dv <= wp.a_max * 3.6 * dt[None, :] * (1 + 1e-12) + 1e-12)[~at_end[:-1]].all() and (H["velocity"] >= 0).all()
    assert (H["velocity"][1:][at_end[:-1]] == 0).all()
    step = np.diff(ids, axis=0)
    assert (step >= -wp.loc_back).all() and (step <= wp.loc_fwd).all()              # every new index lies in the window of the old one
    assert (np.diff(ego_id, axis=0)[lane[1:] == lane[:-1]] >= -1).all()            # no driving backwards on a lane
    # the ego sits within half a lane (+ tracking slack) of the centre line of the lane it is localised on
    m = the_map
    gl = m.road_lane_base[H["road_num"].astype(int) - 1] + lane - 1
    k = m.lane_pt_off[gl] + ego_id
    d = np.hypot(H["x"] - m.x[k], H["y"] - m.y[k])
    assert np.percentile(d, 99) < 2.2 and d.max() < 4.0
    # agents: moved by v * dt along their lane (arclength = index advance x spacing on the 0.5 m map), lateral offset kept
    ag0, ag1 = w.agents, a["agents"]
    T = float(cycles) * dt[:, None]
    adv = (ag1["i"] - ag0["i"]) * 0.5 + (ag1["u"] - ag0["u"])
    free = ag1["i"] < 1990                                                           # (not stopped at the lane end)
    assert np.abs(adv - ag0["v"] * T)[free].max() < 0.05 * cycles                    # arc / S-curve lanes are not exactly 0.5 m apart
    assert (ag1["lane"] == ag0["lane"]).all() and (ag1["lat"] == ag0["lat"]).all() and (ag1["v"] == ag0["v"]).all()
    gla = ag1["lane"].astype(int)
    ka = m.lane_pt_off[gla] + ag1["i"]
    hr = np.radians(m.dir[ka])
    off_left = -(a["ox"] - m.x[ka]) * np.sin(hr) + (a["oy"] - m.y[ka]) * np.cos(hr)
    assert np.abs(off_left - ag1["lat"]).max() < 0.05                                # LEFT-positive offset from the lane centre line


# ---------------------------------------------------------------- CPU: frames vs the reference's own output objects
@pytest.mark.parametrize("kind,seed0,n,cycles", [("highway", 0, 256, 25), ("junction", 70000, 64, 80)])
def test_frames_equal_what_the_reference_publishes(oracle, reference, the_map, kind, seed0, n, cycles):
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(seed0, seed0 + n), cycles=cycles, kind=kind)
    H, OX, OY = ep.all_cycles()
    b = reference.run_with_frames(H, OX, OY)
    a = oracle.run(H, OX, OY, exhaustive=False, threads=4)
    clean = a["trace"]["ub_hits"] == 0
    ctrl, status = oracle.pack_frames(a["rec"].reshape(-1), a["path_xy"].reshape(-1, 2, 200))
    ctrl, status = ctrl.reshape(cycles, n), status.reshape(cycles, n)
    assert (ctrl.view("u1").reshape(cycles, n, -1) == b["ctrl"].view("u1").reshape(cycles, n, -1)).all(axis=2)[clean].all()
    assert (status.view("u1").reshape(cycles, n, -1) == b["status"].view("u1").reshape(cycles, n, -1)).all(axis=2)[clean].all()
    assert b["ctrl"]["sstop"].all() and set(np.unique(b["ctrl"]["desacc_vd"]).tolist()) == {0, 1}
    assert len(np.unique(b["ctrl"]["cnt"])) > 20


def test_frame_layouts(oracle):
    from dmpp_b200 import abi
    import ctypes as C
    assert oracle.lib.oracle_sizeof(7) == abi.ctrl_frame.itemsize == 1656
    assert oracle.lib.oracle_sizeof(8) == abi.status_frame.itemsize == 1632
    assert oracle.lib.oracle_sizeof(9) == abi.agent.itemsize == 32
    assert oracle.lib.oracle_sizeof(10) == C.sizeof(abi.WorldParams)
    assert abi.ctrl_frame.fields["brakedis"][1] == 8 and abi.ctrl_frame.fields["pnts"][1] == 56
    assert abi.status_frame.fields["path_points"][1] == 32


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def planner(the_map):
    from dmpp_b200.planner import Planner
    p = Planner(max_scenes=4096, max_obs=10)
    p.upload_map(the_map)
    yield p
    p.close()


def check_closed_loop(got, want, what):
    assert_records_equal(got["rec"], want["rec"], GPU_REC + ["path_dir_err"], close=DIR_ERR_TOL, what=what + " record")
    assert_records_equal(got["hdr_log"], want["hdr_log"], HDR, what=what + " logged header")
    used = np.arange(got["ox"].shape[1])[None, :] < want["hdr"]["n_obs"][:, None]     # (only the first n_obs points of a scene are written)
    for k in ("obs_log_x", "obs_log_y"):
        assert (same(got[k], want[k]) | ~used[None]).all(), what + " " + k
    assert got["hdr"].tobytes() == want["hdr"].tobytes(), what + " final header"
    assert got["agents"].tobytes() == want["agents"].tobytes(), what + " final agents"
    assert ((same(got["ox"], want["ox"]) & same(got["oy"], want["oy"])) | ~used).all(), what + " final obstacle points"
    assert_records_equal(got["carry"], want["carry"], CARRY + ["plan_count"], what=what + " carry")
    assert same(got["last_path"], want["last_path"]).all(), what + " last_Bpoints"


@pytest.mark.gpu
def test_gpu_closed_loop_equals_oracle(planner, oracle, the_map):
    """4096 worlds x 60 cycles in ONE graph launch, replayed three times from the same initial world"""
    w = world(the_map, 0, 4096)
    want = oracle.run_closed_loop(w.hdr, w.agents, 60, threads=8)
    got = planner.run_closed_loop(w.hdr, w.agents, 60, repeat=3)
    assert got["graph"], "the episode did not run as a CUDA graph"
    check_closed_loop(got, want, "graph")


@pytest.mark.gpu
def test_gpu_closed_loop_direct_launches_and_group_kernel(oracle, the_map):
    """the same launches enqueued directly (DP_EPISODE_GRAPH=0), and the phase-synchronous group kernel inside the loop"""
    from dmpp_b200.planner import Planner
    w = world(the_map, 50000, 512)
    want = oracle.run_closed_loop(w.hdr, w.agents, 40, threads=8)
    for env in ({"DP_EPISODE_GRAPH": "0"}, {"DP_KERNEL": "group"}):
        os.environ.update(env)
        try:
            p = Planner(max_scenes=512, max_obs=10)
        finally:
            for k in env:
                del os.environ[k]
        p.upload_map(the_map)
        got = p.run_closed_loop(w.hdr, w.agents, 40)
        assert not got["graph"]
        check_closed_loop(got, want, str(env))
        p.close()


@pytest.mark.gpu
def test_gpu_closed_loop_log_replays_open_loop(planner, the_map):
    """the inputs every closed-loop cycle ran on, fed back through the scripted (open-loop) path, give the same records"""
    w = world(the_map, 9000, 1024)
    got = planner.run_closed_loop(w.hdr, w.agents, 30)
    rep = planner.run_episodes(got["hdr_log"], got["obs_log_x"], got["obs_log_y"], trace=False, paths=False)
    assert rep["rec"].tobytes() == got["rec"].tobytes()


@pytest.mark.gpu
def test_gpu_world_step_ragged_and_windows(planner, oracle, the_map):
    """single steps: ragged obstacle counts, an ego at the lane end, agents running off their lanes, other window sizes"""
    from dmpp_b200 import abi
    w = world(the_map, 123, 600)
    rng = np.random.default_rng(5)
    w.hdr["n_obs"] = rng.integers(0, 11, 600)
    w.agents["i"][::7] = 1996
    w.agents["v"][::7] = 40.0
    wp = planner.world_params()
    wp.loc_back, wp.loc_fwd, wp.a_max = 3, 17, 1.5
    planner.set_world_params(wp)
    try:
        want = oracle.run_closed_loop(w.hdr, w.agents, 12, wp=wp, threads=4)
        got = planner.run_closed_loop(w.hdr, w.agents, 12)
        check_closed_loop(got, want, "ragged")
    finally:
        planner.set_world_params(planner.world_params())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed0,n,cycles", [("highway", 0, 1024, 25), ("junction", 70000, 128, 60)])
def test_gpu_frames_equal_oracle(planner, oracle, the_map, kind, seed0, n, cycles):
    from dmpp_b200 import scenes
    ep = scenes.Episodes(the_map, np.arange(seed0, seed0 + n), cycles=cycles, kind=kind)
    H, OX, OY = ep.all_cycles()
    want = oracle.run(H, OX, OY, exhaustive=False, threads=8)
    planner.reset(0, n)
    for c in range(cycles):
        o = planner.cycle(np.ascontiguousarray(H[c]), np.ascontiguousarray(OX[c]), np.ascontiguousarray(OY[c]))
        ctrl, status = planner.pack_frames(o["rec"])
        wc, ws = oracle.pack_frames(want["rec"][c], want["path_xy"][c])
        assert ctrl.tobytes() == wc.tobytes(), "PlanningOut frame, cycle %d" % c
        assert status.tobytes() == ws.tobytes(), "PlanningStatus frame, cycle %d" % c
    only_ctrl, none = planner.pack_frames(o["rec"], status=False)
    assert none is None and only_ctrl.tobytes() == ctrl.tobytes()
